"""Multi-GPU sharding of the env batch: one process per GPU (torch.distributed, NCCL on B200 / gloo in CPU tests).

The step has no cross-env term, so ranks never exchange data on the step path: rank r owns the
contiguous slice [offset, offset+count) of the global env index space and keys its RNG streams by GLOBAL env index
(`env_index_offset`), which makes results independent of the world size.  Collectives exist only for consumers that want
global views: scalar statistics (all-reduce) or, optionally, the gathered observation batch (all-gather).
"""
import os


def shard_envs(total_envs, rank, world_size):
    """Contiguous, balanced partition: returns (count, offset) for `rank`."""
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return count, offset


def rank_world():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_sharded(env_id, total_envs, **kwargs):
    """`make` for this rank's shard of a global batch of `total_envs` environments (device = LOCAL_RANK by default)."""
    from .envs import make

    rank, world = rank_world()
    count, offset = shard_envs(total_envs, rank, world)
    kwargs.setdefault("device", int(os.environ.get("LOCAL_RANK", "0")))
    return make(env_id, num_envs=count, env_index_offset=offset, **kwargs)


def global_stats(reward, terminated):
    """(mean reward, number of terminated envs) over ALL ranks: one all-reduce of two scalars."""
    import torch
    import torch.distributed as dist

    t = torch.stack([reward.double().sum(), terminated.double().sum(), torch.tensor(float(reward.numel()), device=reward.device, dtype=torch.float64)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return float(t[0] / t[2]), int(t[1].item())


class PackedStepOutputs:
    """The ONE collective of the sharded layout (BASELINE.json north_star): every rank lets the step kernel write its
    observation / next reference / reward / terminated directly into ONE contiguous byte buffer (the C-ABI's output tensors are
    caller-owned), so that a single `all_gather_into_tensor` returns the aggregated batch of all ranks.

        out = PackedStepOutputs(n_local, n_state, n_ref, dtype, device)
        env.sim.bind_outputs(*out.local_views())          # the kernel writes straight into the packed buffer
        env.step(actions); obs, ref, rew, term = out.gather()   # [world * n_local, ...] views, rank-major

    Sections are 16-byte aligned (the kernel's vector stores need it).  Same n_local on every rank."""

    def __init__(self, n_local, n_state, n_ref, dtype, device):
        import torch
        import torch.distributed as dist

        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.n, self.n_state, self.n_ref, self.dtype = int(n_local), int(n_state), int(n_ref), dtype
        isz = torch.empty((), dtype=dtype).element_size()
        sizes = [self.n * n_state * isz, self.n * n_ref * isz, self.n * isz, self.n]
        self.offsets, off = [], 0
        for sz in sizes:
            self.offsets.append(off)
            off += (sz + 15) // 16 * 16
        self.nbytes = off
        self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(self.world * self.nbytes, dtype=torch.uint8, device=device) if self.world > 1 else self.local

    def _views(self, buf, base):
        import torch

        o = self.offsets
        isz = torch.empty((), dtype=self.dtype).element_size()
        obs = buf[base + o[0]: base + o[0] + self.n * self.n_state * isz].view(self.dtype).view(self.n, self.n_state)
        ref = buf[base + o[1]: base + o[1] + self.n * self.n_ref * isz].view(self.dtype).view(self.n, self.n_ref)
        rew = buf[base + o[2]: base + o[2] + self.n * isz].view(self.dtype)
        term = buf[base + o[3]: base + o[3] + self.n]
        return obs, ref, rew, term

    def local_views(self):
        return self._views(self.local, 0)

    def gather(self):
        """one NCCL (or gloo) all-gather of the packed buffer; returns per-field tensors [world, n_local, ...] (rank-major views)"""
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.local)
        per = [self._views(self.gathered, r * self.nbytes) for r in range(self.world)]
        return tuple(torch.stack([p[k] for p in per]) if self.world > 1 else per[0][k].unsqueeze(0) for k in range(4))

    def gather_raw(self):
        """the collective alone (no per-field re-assembly): what a consumer that reads the rank-major sections directly pays"""
        import torch.distributed as dist

        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.local)
        return self.gathered


class OverlappedGather:
    """Double-buffered `PackedStepOutputs`: step k writes into buffer k % 2 on the caller's stream, its all-gather runs on a side
    stream while step k+1 computes; buffer b is reused only after the gather that read it has finished (event, no host sync).

        og = OverlappedGather(env.sim, torch.float32)
        for a in actions: b = og.step(a)          # returns the buffer index whose gather was just enqueued
        og.finish()                               # caller's stream waits for the outstanding gathers
        obs, ref, rew, term = og.views(b)         # rank-major [world, n_local, ...] views of buffer b
    """

    def __init__(self, sim, dtype):
        import torch

        self.sim, self.torch = sim, torch
        self.bufs = [PackedStepOutputs(sim.n, sim.n_state, sim.n_ref, dtype, sim.device) for _ in range(2)]
        self.nbytes = self.bufs[0].nbytes
        self.side = torch.cuda.Stream(device=sim.device)
        self.done = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self._saved = (sim._reuse, sim._out, getattr(sim, "_out_ptrs", None))
        self._bound = []
        for b in self.bufs:  # resolve the output pointers once per buffer
            sim.bind_outputs(*b.local_views())
            self._bound.append(sim._out)

    def step(self, action):
        torch, b = self.torch, self.k % 2
        cur = torch.cuda.current_stream(self.sim.device)
        if self.done[b] is not None:
            cur.wait_event(self.done[b])
        self.sim._out, self.sim._out_ptrs = self._bound[b], None
        self.sim.step(action)
        self.ready[b].record(cur)
        self.side.wait_event(self.ready[b])
        with torch.cuda.stream(self.side):
            self.bufs[b].gather_raw()
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[b] = ev
        self.k += 1
        return b

    def finish(self):
        cur = self.torch.cuda.current_stream(self.sim.device)
        for ev in self.done:
            if ev is not None:
                cur.wait_event(ev)

    def views(self, b):
        buf = self.bufs[b]
        per = [buf._views(buf.gathered, r * buf.nbytes) for r in range(buf.world)]
        return tuple(self.torch.stack([p[k] for p in per]) for k in range(4))

    def release(self):
        self.finish()
        self.sim._reuse, self.sim._out, self.sim._out_ptrs = self._saved


def all_gather_batch(*tensors):
    """Optional single all-gather of per-rank [n_local, ...] tensors into global [N, ...] tensors (same n_local on every
    rank).  For PMSM at N=2^20 this moves ~72 MB per step — several times the step itself; data-parallel
    learners should keep observations rank-local and use global_stats instead."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors
    out = []
    for t in tensors:
        g = torch.empty((dist.get_world_size() * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(g, t.contiguous())
        out.append(g)
    return tuple(out)
