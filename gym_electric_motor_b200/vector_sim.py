"""Low-level handle wrapper: one `VectorSim` = one gemb200_handle = N envs of one motor/converter/load/solver
combination on one CUDA device.  Tensors are torch tensors on that device (torch is plumbing: memory + streams);
all compute happens in libgemb200.so through the C-ABI (include/gemb200.h).

This is the batched counterpart of the reference's `SCMLSystem` + the per-step part of `ElectricMotorEnvironment`
(physical_systems.py:13-287, core.py:300-371).
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi as K


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class VectorSim:
    def __init__(self, cfg, reuse_outputs=True):
        self._lib = K.load_library()
        if not torch.cuda.is_available():
            raise K.GemB200Error("no CUDA device visible: the B200 path has no CPU fallback")
        self.cfg = cfg
        self.device = torch.device("cuda", int(cfg.device))
        d = [C.c_int32() for _ in range(4)]
        K.check(self._lib.gemb200_query_dims(C.byref(cfg), *[C.byref(x) for x in d]), "gemb200_query_dims")
        self.n_state, self.n_ode, self.n_act, self.n_ref = [x.value for x in d]
        self.n = int(cfg.n_envs)
        self.finite = bool(cfg.finite)
        self.soa = cfg.layout == K.LAYOUT_SOA
        self.dtype = torch.float32 if cfg.dtype == K.F32 else torch.float64
        self.np_dtype = np.float32 if cfg.dtype == K.F32 else np.float64
        self.act_dtype = torch.int32 if self.finite else self.dtype
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            K.check(self._lib.gemb200_create(C.byref(cfg), C.byref(h)), "gemb200_create")
        self._h = h
        self._reuse = reuse_outputs
        self._out = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gemb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _shape(self, k):
        return (k, self.n) if self.soa else (self.n, k)

    def _alloc_outputs(self):
        if self._reuse and self._out is not None:
            return self._out
        self._out_ptrs = None
        out = (
            torch.empty(self._shape(self.n_state), dtype=self.dtype, device=self.device),
            torch.empty(self._shape(self.n_ref), dtype=self.dtype, device=self.device),
            torch.empty(self.n, dtype=self.dtype, device=self.device),
            torch.empty(self.n, dtype=torch.uint8, device=self.device),
        )
        if self._reuse:
            self._out = out
        return out

    def bind_outputs(self, obs, ref, reward, terminated):
        """Let the step / reset launches write into caller-owned tensors (e.g. the sections of a packed all-gather buffer,
        distributed.PackedStepOutputs).  Shapes, dtypes and device must match what `step` returns."""
        exp = (self._shape(self.n_state), self._shape(self.n_ref), (self.n,), (self.n,))
        for t, shp, dt in zip((obs, ref, reward, terminated), exp, (self.dtype, self.dtype, self.dtype, torch.uint8)):
            if tuple(t.shape) != tuple(shp) or t.dtype != dt or t.device != self.device or not t.is_contiguous():
                raise ValueError(f"output tensor mismatch: need {tuple(shp)} {dt} contiguous on {self.device}, got {tuple(t.shape)} {t.dtype} on {t.device}")
            if t.data_ptr() % 16:
                raise ValueError("output tensors must be 16-byte aligned")
        self._reuse, self._out, self._out_ptrs = True, (obs, ref, reward, terminated), None

    def _as_action(self, action):
        if isinstance(action, torch.Tensor) and action.dtype == self.act_dtype and action.device == self.device and action.is_contiguous() \
                and action.numel() == self.n * self.n_act:
            return action  # fast path: nothing to convert (the launch only needs the pointer)
        a = torch.as_tensor(action, device=self.device)
        if a.dtype != self.act_dtype:
            a = a.to(self.act_dtype)
        a = a.reshape(self._shape(self.n_act))
        return a.contiguous()

    # ------------------------------------------------------------------ env API (device tensors)
    def reset(self, mask=None):
        """env.reset for all (mask=None) or the masked envs; returns (obs, ref_next) device tensors.
        With a mask, rows of unmasked envs keep their previous content."""
        obs, ref, _, _ = self._alloc_outputs()
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        K.check(self._lib.gemb200_reset(self._h, _ptr(m), _ptr(obs), _ptr(ref) if self.n_ref else None, self._stream()), "gemb200_reset")
        return obs, ref

    def reseed(self, seed):
        """re-key the RNG streams and start the handle over (gemb200_reseed): equal seeds -> identical episodes"""
        K.check(self._lib.gemb200_reseed(self._h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), self._stream()), "gemb200_reseed")
        self.cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF

    def set_device_clock(self, enable=True):
        """Device-resident clock (gemb200_set_device_clock): while on, step / rollout / reset launches read the RNG call id, the step count
        and the dead-time ring position from device memory and advance them with a one-thread kernel, so the launches can be captured in a
        CUDA graph and replayed (graph.CapturedSteps).  Same results as with the host clock, bit for bit."""
        K.check(self._lib.gemb200_set_device_clock(self._h, 1 if enable else 0, self._stream()), "gemb200_set_device_clock")

    def clock(self):
        """(number of API calls that drew random numbers, number of env steps) so far; synchronises when the device clock is on"""
        a, b = C.c_uint64(), C.c_uint64()
        K.check(self._lib.gemb200_get_clock(self._h, C.byref(a), C.byref(b), self._stream()), "gemb200_get_clock")
        return int(a.value), int(b.value)

    def set_env_params(self, motor_param=None, load_param=None):
        """Per-env parameter blocks (gemb200_set_env_params): motor_param [N, 16] / load_param [N, 8] float64 host arrays in the slot order of
        `_cabi.MP_*` / `_cabi.LP_*`; None keeps the configuration's values; both None: back to the shared coefficients."""
        mp = None if motor_param is None else np.ascontiguousarray(motor_param, dtype=np.float64).reshape(self.n, K.MAX_MOTOR_PARAM)
        lp = None if load_param is None else np.ascontiguousarray(load_param, dtype=np.float64).reshape(self.n, 8)
        vp = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)  # noqa: E731
        K.check(self._lib.gemb200_set_env_params(self._h, vp(mp), vp(lp)), "gemb200_set_env_params")

    def step(self, action):
        """env.step: returns (obs, ref_next, reward, terminated) device tensors (views of reused buffers unless
        reuse_outputs=False)."""
        a = self._as_action(action)
        out = self._alloc_outputs()
        if self._reuse:
            if getattr(self, "_out_ptrs", None) is None:
                obs, ref, rew, term = out
                self._out_ptrs = (_ptr(obs), _ptr(ref) if self.n_ref else None, _ptr(rew), _ptr(term))
            po, pr, pw, pt = self._out_ptrs
        else:
            obs, ref, rew, term = out
            po, pr, pw, pt = _ptr(obs), _ptr(ref) if self.n_ref else None, _ptr(rew), _ptr(term)
        rc = self._lib.gemb200_step(self._h, a.data_ptr(), po, pr, pw, pt, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            K.check(rc, "gemb200_step")
        return out

    def rollout(self, actions, record_every=0):
        """K open-loop env.step calls fused into ONE launch (gemb200_rollout_record): every env's record stays in registers for all K
        steps; bit-identical to K calls of `step`.  actions: [K, N, n_act] (SoA layout: [K, n_act, N]).
        record_every = 0 -> the outputs of the last step (obs, ref, reward, terminated), shapes as `step`;
        record_every = m >= 1 -> the outputs of steps m, 2m, ... stacked on a leading axis of length K // m (m = 1: full trajectory)."""
        a = actions if (isinstance(actions, torch.Tensor) and actions.dtype == self.act_dtype and actions.device == self.device and actions.is_contiguous()) \
            else torch.as_tensor(actions, device=self.device).to(self.act_dtype).contiguous()
        k = int(a.shape[0])
        if a.numel() != k * self.n * self.n_act:
            raise ValueError(f"actions must hold K x {self.n} x {self.n_act} values")
        m = int(record_every)
        if m == 0:
            obs, ref, rew, term = self._alloc_outputs()
        else:
            s = k // m
            obs = torch.empty((s,) + self._shape(self.n_state), dtype=self.dtype, device=self.device)
            ref = torch.empty((s,) + self._shape(self.n_ref), dtype=self.dtype, device=self.device)
            rew = torch.empty((s, self.n), dtype=self.dtype, device=self.device)
            term = torch.empty((s, self.n), dtype=torch.uint8, device=self.device)
        K.check(self._lib.gemb200_rollout_record(self._h, _ptr(a), k, m, _ptr(obs), _ptr(ref) if self.n_ref else None, _ptr(rew), _ptr(term), self._stream()),
                "gemb200_rollout_record")
        return obs, ref, rew, term

    def rollout_into(self, actions, n_steps, record_every, obs, ref, rew, term):
        """Raw variant for benchmarking: caller-owned output tensors (any may be None), no allocation, no conversion."""
        K.check(self._lib.gemb200_rollout_record(self._h, _ptr(actions), int(n_steps), int(record_every), _ptr(obs), _ptr(ref) if self.n_ref else None,
                                                 _ptr(rew), _ptr(term), self._stream()), "gemb200_rollout_record")

    # ------------------------------------------------------------------ host-buffer API (numpy)
    def step_host(self, action, out=None):
        """Same step through HOST buffers (the C-ABI does H2D, launch, D2H, sync).  `out` = tuple of numpy arrays to
        fill (obs, ref, reward, terminated); allocated when None."""
        a = np.ascontiguousarray(action, dtype=np.int32 if self.finite else self.np_dtype).reshape(self._shape(self.n_act))
        if out is None:
            out = (np.empty(self._shape(self.n_state), self.np_dtype), np.empty(self._shape(self.n_ref), self.np_dtype),
                   np.empty(self.n, self.np_dtype), np.empty(self.n, np.uint8))
        obs, ref, rew, term = out
        vp = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
        K.check(self._lib.gemb200_step_host(self._h, vp(a), vp(obs), vp(ref) if self.n_ref else None, vp(rew), vp(term)), "gemb200_step_host")
        return out

    def step_host_ptr(self, a_ptr, obs_ptr, ref_ptr, rew_ptr, term_ptr):
        """Raw-pointer variant for pinned torch host tensors (bench e2e leg)."""
        K.check(self._lib.gemb200_step_host(self._h, a_ptr, obs_ptr, ref_ptr, rew_ptr, term_ptr), "gemb200_step_host")

    def reset_host(self, mask=None):
        obs = np.zeros(self._shape(self.n_state), self.np_dtype)
        ref = np.zeros(self._shape(self.n_ref), self.np_dtype)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        vp = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)  # noqa: E731
        K.check(self._lib.gemb200_reset_host(self._h, vp(m), vp(obs), vp(ref) if self.n_ref else None), "gemb200_reset_host")
        return obs, ref

    # ------------------------------------------------------------------ state access
    def get_ode_state(self):
        out = torch.empty((self.n, self.n_ode), dtype=torch.float64, device=self.device)
        K.check(self._lib.gemb200_get_ode_state(self._h, _ptr(out), self._stream()), "gemb200_get_ode_state")
        return out

    def set_ode_state(self, y):
        y = torch.as_tensor(y, dtype=torch.float64, device=self.device).reshape(self.n, self.n_ode).contiguous()
        K.check(self._lib.gemb200_set_ode_state(self._h, _ptr(y), self._stream()), "gemb200_set_ode_state")
        torch.cuda.current_stream(self.device).synchronize()  # y may be a temporary

    def get_reference(self):
        out = torch.empty((self.n, self.n_ref), dtype=torch.float64, device=self.device)
        if self.n_ref:
            K.check(self._lib.gemb200_get_reference(self._h, _ptr(out), self._stream()), "gemb200_get_reference")
        return out

    def set_reference(self, r):
        if not self.n_ref:
            return
        r = torch.as_tensor(r, dtype=torch.float64, device=self.device).reshape(self.n, self.n_ref).contiguous()
        K.check(self._lib.gemb200_set_reference(self._h, _ptr(r), self._stream()), "gemb200_set_reference")
        torch.cuda.current_stream(self.device).synchronize()

    def state_dict(self):
        size = self._lib.gemb200_checkpoint_size(self._h)
        buf = np.empty(size, dtype=np.uint8)
        K.check(self._lib.gemb200_checkpoint_save(self._h, buf.ctypes.data_as(C.c_void_p)), "gemb200_checkpoint_save")
        return {"blob": buf}

    def load_state_dict(self, sd):
        buf = np.ascontiguousarray(sd["blob"], dtype=np.uint8)
        if buf.size != self._lib.gemb200_checkpoint_size(self._h):
            raise ValueError("checkpoint size mismatch")
        K.check(self._lib.gemb200_checkpoint_load(self._h, buf.ctypes.data_as(C.c_void_p)), "gemb200_checkpoint_load")

    # ------------------------------------------------------------------ measurement helpers
    @property
    def launch_count(self):
        return int(self._lib.gemb200_launch_count(self._h))

    def time_begin(self):
        K.check(self._lib.gemb200_kernel_time_begin(self._h, self._stream()))

    def time_end(self):
        ms = C.c_float()
        K.check(self._lib.gemb200_kernel_time_end(self._h, self._stream(), C.byref(ms)))
        return ms.value
