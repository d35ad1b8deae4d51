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


class _RawDeviceBytes:
    """zero-copy torch view of library-allocated device memory (CUDA array interface)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGather:
    """The sharded layout's ONE collective fused into the step itself (include/gemb200.h: gemb200_bind_peers): every rank's step kernel stores
    obs | ref | reward | terminated of its shard straight into section `rank` of EVERY rank's gather buffer over NVLink (peer stores through
    CUDA-IPC mappings of library-allocated buffers), so no all-gather runs at all; a per-(buffer, source) flag written after the step replaces the
    collective's synchronisation and a credit flag guards the reuse of the two buffers.

        pg = PeerGather(env.sim, torch.float32)         # collective: exchanges the IPC handles (torch.distributed, any backend)
        for a in actions: b = pg.step(a)                # step k on the caller's stream; its arrival is awaited on a side stream
        pg.finish()                                     # caller's stream waits until every outstanding step has arrived from every rank
        obs, ref, rew, term = pg.views(b)               # [world, n_local, ...] rank-major views of buffer b (valid until 2 more steps)
    """

    FLAG_BYTES = 4096

    def __init__(self, sim, dtype):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _cabi as K

        self.sim, self.torch, self._lib, self._C = sim, torch, K.load_library(), C
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if self.world > 8:
            raise ValueError("PeerGather: at most 8 ranks (one NVLink domain)")
        self.layout = PackedStepOutputs.__new__(PackedStepOutputs)  # section layout only (no buffers)
        isz = torch.empty((), dtype=dtype).element_size()
        lay = self.layout
        lay.world, lay.n, lay.n_state, lay.n_ref, lay.dtype = self.world, sim.n, sim.n_state, sim.n_ref, dtype
        sizes = [sim.n * sim.n_state * isz, sim.n * sim.n_ref * isz, sim.n * isz, sim.n]
        lay.offsets, off = [], 0
        for sz in sizes:
            lay.offsets.append(off)
            off += (sz + 15) // 16 * 16
        lay.nbytes = self.nbytes = off
        self.dev = int(sim.device.index)
        self.total = 2 * self.world * self.nbytes + self.FLAG_BYTES
        ptr, handle = C.c_void_p(), (C.c_char * 64)()
        K.check(self._lib.gemb200_peer_buffer_alloc(self.dev, self.total, C.byref(ptr), handle), "gemb200_peer_buffer_alloc")
        self.own = ptr.value
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(handle.raw))
        self.bases, self._opened = [], []
        for r in range(self.world):
            if r == self.rank:
                self.bases.append(self.own)
                continue
            m = C.c_void_p()
            hb = (C.c_char * 64).from_buffer_copy(handles[r])
            K.check(self._lib.gemb200_peer_buffer_open(self.dev, hb, C.byref(m)), "gemb200_peer_buffer_open")
            self.bases.append(m.value)
            self._opened.append(m.value)
        self.buf = torch.as_tensor(_RawDeviceBytes(self.own, self.total), device=sim.device)
        flag0 = 2 * self.world * self.nbytes
        # flag pointer tables (device): for buffer b, the address of ready[b][rank] / credit[b][rank] in every rank's buffer
        def table(kind, b):
            off_ = flag0 + kind * 1024 + (b * self.world + self.rank) * 4
            return torch.tensor([base + off_ for base in self.bases], dtype=torch.int64, device=sim.device)

        self._ready_tab = [table(0, b) for b in range(2)]
        self._credit_tab = [table(1, b) for b in range(2)]
        self._ready_local = [self.own + flag0 + b * self.world * 4 for b in range(2)]
        self._credit_local = [self.own + flag0 + 1024 + b * self.world * 4 for b in range(2)]
        self.err = torch.zeros(1, dtype=torch.int32, device=sim.device)
        self.side = torch.cuda.Stream(device=sim.device)
        self._arrived = [None, None]
        self.k = 0
        self._saved = (sim._reuse, sim._out, getattr(sim, "_out_ptrs", None))
        self._local = []
        for b in range(2):
            base = (b * self.world + self.rank) * self.nbytes
            self._local.append(lay._views(self.buf, base))
        delta = (C.c_int64 * self.world)(*[base - self.own for base in self.bases])
        K.check(self._lib.gemb200_bind_peers(sim._h, self.world, delta), "gemb200_bind_peers")
        if self.world > 1:
            dist.barrier()  # every rank has mapped every buffer before the first peer store

    def _vp(self, x):
        return self._C.c_void_p(int(x))

    def step(self, action):
        torch, K_ = self.torch, self.k + 1
        b = self.k % 2
        sim, cur = self.sim, torch.cuda.current_stream(self.sim.device)
        from . import _cabi as K

        if K_ > 2:  # buffer b still holds step K_ - 2: every rank must have consumed it
            K.check(self._lib.gemb200_peer_wait(sim._h, self.world, self._vp(self._credit_local[b]), K_ - 2, self._vp(self.err.data_ptr()), self._vp(cur.cuda_stream)),
                    "gemb200_peer_wait")
        sim._reuse, sim._out, sim._out_ptrs = True, self._local[b], None
        sim.step(action)
        K.check(self._lib.gemb200_peer_signal(sim._h, self.world, self._vp(self._ready_tab[b].data_ptr()), K_, self._vp(cur.cuda_stream)), "gemb200_peer_signal")
        # consumer side: await the arrival of step K_ from every rank, then hand the buffer back
        with torch.cuda.stream(self.side):
            K.check(self._lib.gemb200_peer_wait(sim._h, self.world, self._vp(self._ready_local[b]), K_, self._vp(self.err.data_ptr()), self._vp(self.side.cuda_stream)),
                    "gemb200_peer_wait")
            ev = torch.cuda.Event()
            ev.record(self.side)
            K.check(self._lib.gemb200_peer_signal(sim._h, self.world, self._vp(self._credit_tab[b].data_ptr()), K_, self._vp(self.side.cuda_stream)), "gemb200_peer_signal")
        self._arrived[b] = ev
        self.k += 1
        return b

    def finish(self):
        cur = self.torch.cuda.current_stream(self.sim.device)
        for ev in self._arrived:
            if ev is not None:
                cur.wait_event(ev)

    def check(self):
        """host-side check of the flag protocol's time-out indicator (synchronises)"""
        self.torch.cuda.synchronize(self.sim.device)
        e = int(self.err.item())
        if e:
            raise RuntimeError(f"PeerGather: rank {self.rank} gave up waiting for the flag of rank {e - 1}")

    def views(self, b):
        per = [self.layout._views(self.buf, (b * self.world + r) * self.nbytes) for r in range(self.world)]
        return tuple(self.torch.stack([p[k] for p in per]) for k in range(4))

    def release(self):
        import torch.distributed as dist

        from . import _cabi as K

        self.finish()
        self.torch.cuda.synchronize(self.sim.device)
        K.check(self._lib.gemb200_bind_peers(self.sim._h, 0, None), "gemb200_bind_peers")
        self.sim._reuse, self.sim._out, self.sim._out_ptrs = self._saved
        if self.world > 1:
            dist.barrier()  # nobody stores into a buffer that is about to be unmapped
        for m in self._opened:
            self._lib.gemb200_peer_buffer_close(self.dev, self._vp(m))
        self._opened = []
        if self.world > 1:
            dist.barrier()
        del self.buf
        self._lib.gemb200_peer_buffer_free(self.dev, self._vp(self.own))


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
