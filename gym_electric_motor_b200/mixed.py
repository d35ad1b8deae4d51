"""Mixed-motor batches (BASELINE.json configs[4]: PMSM + SynRM + EESM interleaved, per-type kernel dispatch).

A kernel instantiation is specialised per motor family, so a heterogeneous batch is stored PHYSICALLY SEGMENTED per type
(one handle = one contiguous SoA block per type) while the caller may think in interleaved global ids (env g has type
g mod T).  `step` issues one launch per type, each on its own CUDA stream, so the launches overlap on the device; the
caller's stream waits on all of them (events), no host synchronisation.
"""
import torch

from .envs import make


class MixedEnvBatch:
    """`env_ids`: list of T env ids (or (env_id, kwargs) tuples); `num_envs`: TOTAL number of envs, split evenly by type in
    interleaved order (global env g -> type g % T, local index g // T)."""

    def __init__(self, env_ids, num_envs, device=0, **common_kwargs):
        self.specs = [(e, {}) if isinstance(e, str) else (e[0], dict(e[1])) for e in env_ids]
        T = len(self.specs)
        if num_envs % T:
            raise ValueError("num_envs must be a multiple of the number of types")
        self.num_types, self.num_envs, self.per_type = T, int(num_envs), int(num_envs) // T
        offset = int(common_kwargs.pop("env_index_offset", 0))
        self.envs = []
        for t, (env_id, kw) in enumerate(self.specs):
            k = dict(common_kwargs)
            k.update(kw)
            # global index space: type t owns [offset + t*per_type, offset + (t+1)*per_type) of the RNG key space
            self.envs.append(make(env_id, num_envs=self.per_type, device=device, env_index_offset=offset + t * self.per_type, **k))
        self.device = torch.device("cuda", int(device) if not isinstance(device, str) else int(str(device).split(":")[-1]))
        self._streams = None
        self._events = None

    def _ensure_streams(self):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in self.envs]
            self._events = [torch.cuda.Event() for _ in self.envs]
            self._start = torch.cuda.Event()

    # interleaved <-> segmented helpers -------------------------------------------------------------------------
    def split_interleaved(self, x):
        """[N, ...] in interleaved global order -> list of T tensors [N/T, ...] (strided views, no copy)."""
        return [x[t :: self.num_types] for t in range(self.num_types)]

    def reset(self, seed=None):
        return [env.reset(seed=seed) for env in self.envs]

    def step(self, actions):
        """`actions`: list of T action tensors (per type, [N/T, n_act_t]).  Returns the list of per-type step results.
        Per-type launches run concurrently on separate streams."""
        self._ensure_streams()
        cur = torch.cuda.current_stream(self.device)
        self._start.record(cur)
        results = []
        for env, a, s, e in zip(self.envs, actions, self._streams, self._events):
            s.wait_event(self._start)
            with torch.cuda.stream(s):
                results.append(env.step(a))
            e.record(s)
        for e in self._events:
            cur.wait_event(e)
        return results

    def close(self):
        for env in self.envs:
            env.close()
