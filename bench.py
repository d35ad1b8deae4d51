#!/usr/bin/env python
"""bench.py — env-steps/s of the fused GEM step on B200 (BASELINE.json metric), with roofline + CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs-per-gpu E]

Workload (BASELINE.json configs[1]/metric): Cont-CC-PMSM-v0, RK4 (one step per tau = 1e-4), 2^20 envs per GPU, synthetic
U(-1,1)^3 actions, Wiener references, in-kernel auto-reset; weak scaling: every rank steps its own 2^20-env shard
(keyed by global env index), no data-path collective.

One "step" = one batched env.step = ONE launch of step_kernel over the rank's shard.
 * value    : whole-job env-steps/s with actions resident in HBM.  The K timed steps run back to back (one CUDA-event pair
              on the launching stream around all K launches) and rotate over R=4 independent replicas of the 2^20-env batch
              and 8 action tensors, so that every byte a launch touches was last touched >= 3 launches (~500 MB of traffic,
              4x the 126 MB L2) earlier: inputs larger than L2, no flush kernels inside the timed region.
              ms_per_step = event time / K, max over ranks.
 * cold_events : the same step timed one launch at a time (CUDA events around every launch, 256 MiB L2 flush before it);
              includes ~8 us of event/launch overhead per step and is reported for reference.
 * e2e      : same metric through the public host-buffer entry point (gemb200_step_host): pinned host actions -> H2D ->
              launch -> D2H of obs/ref/reward/terminated -> sync, every step.
 * roofline : algorithmic bytes per env-step (SURVEY.md §8d, 129 B for PMSM) * envs / mean kernel time vs the measured
              HBM copy bandwidth in MEASURED_PEAKS.json.
 * cpu_baseline : the float64 C oracle (a port of the reference's algorithm, oracle/gem_oracle.c) on this box's host
              cores, bounded sample.  `--impl reference` runs only that arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENV_ID = "Cont-CC-PMSM-v0"
B_ALG = 129  # algorithmic bytes per env-step, SURVEY.md §8(d): 4*3 act + 2*4*4 state + 4*14 obs + 4*2 ref + 2*4*2 gen + 4 + 1
METRIC = "env-steps/sec at N=2^20 PMSM, 1/2/4/8 GPU; HBM GB/s vs roofline"
UNIT = "env-steps/s"


def make_env(n_envs, device=0, rank=0):
    import gym_electric_motor_b200 as gem

    return gem.make(ENV_ID, num_envs=n_envs, device=device, dtype="float32", ode_solver=gem.physical_systems.RK4Solver(),
                    autoreset="same_step", seed=0, env_index_offset=rank * n_envs)


def workload_config(n_envs, n_gpus):
    return {"workload": f"{ENV_ID} x {n_envs} envs/GPU, RK4 x1 per tau=1e-4, ContB6 + IdealSupply + ConstantSpeedLoad(100 rad/s), "
                        "Wiener refs (i_sd,i_sq) + WSE reward + SquaredConstraint fused, same-step auto-reset",
            "env_id": ENV_ID, "envs_per_gpu": n_envs, "global_envs": n_envs * n_gpus, "solver": "rk4x1", "tau": 1e-4,
            "parallelism": f"env-shard x{n_gpus} (no collective)", "l2": "inputs larger than L2: timed steps rotate over 4 replicas of the env batch (4 x 166 MB touched per cycle), back-to-back launches",
            "layout": "obs [N,14] row-per-env (AoS), state SoA"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("step_kernel_pmsm_f32_aos_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        """Start early (nvidia-smi needs ~100 ms before its first row); rows are stamped on arrival, mark()/stop() bracket the timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.03)  # let the last row of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t_begin = getattr(self, "t_begin", 0.0)
        inside = [r for (t, r) in self.rows if t_begin <= t <= t_end + 0.03]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: take the rows closest to it
            inside = [r for (t, r) in self.rows if t_begin - 0.25 <= t <= t_end + 0.25]
            window = "timed region +-250 ms"
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_arm(n_envs, steps, warmup, budget_s=None):
    """Oracle (C port of the reference algorithm) on the host cores; returns (steps_per_s, cores, sample description)."""
    import numpy as np

    from oracle.gem_oracle import Oracle
    from gym_electric_motor_b200 import _cabi as K

    cores = os.cpu_count() or 1
    env = make_env(n_envs)  # host-side spec only; no device handle is created until the env is used
    cfg = env.build_config()
    cfg.dtype = K.F64
    ora = Oracle(cfg, nthreads=cores)
    ora.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, size=(4, n_envs, 3))
    chunk = 16  # steps per rollout call: the oracle's worker threads persist over a call (two barriers per step, no thread creation)
    for k in range(max(1, warmup // chunk)):
        ora.rollout(acts, chunk)
    t0 = time.perf_counter()
    done = 0
    while True:
        c = chunk if budget_s is not None else min(chunk, steps - done)
        ora.rollout(acts, c)
        done += c
        el = time.perf_counter() - t0
        if budget_s is None:
            if done >= steps:
                break
        elif el >= budget_s or done >= 100000:
            break
    el = time.perf_counter() - t0
    return n_envs * done / el, cores, f"{ENV_ID}, {n_envs} envs x {done} steps, RK4 x1, float64 C oracle, {cores} threads, {el:.1f} s"


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm for the same path on the host cores (oracle port; the Python
    reference itself cannot travel to this box and runs ~1e4 steps/s/core)."""
    if rank != 0:
        return
    n_ref = 65536
    v, cores, sample = cpu_arm(n_ref, max(args.steps, 1), max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n_ref / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(1 << 20, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n = args.envs_per_gpu
    R = 4  # replicas of the env batch that the timed steps rotate over (working set >> L2)
    envs = [make_env(n, device=local_rank, rank=rank * R + r) for r in range(R)]
    env = envs[0]
    sim = env.sim
    for e in envs:
        e.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = [torch.rand((n, 3), generator=gen, device=dev, dtype=torch.float32) * 2 - 1 for _ in range(8)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: K back-to-back launches rotating over R replicas, one event pair ----------------
    for k in range(max(W, R)):
        envs[k % R].step(pool[k % 8])
    barrier()
    if rank == 0:
        sampler.wait_first()
        sampler.mark()
    l0 = sum(e.sim.launch_count for e in envs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for k in range(K):
        envs[k % R].step(pool[k % 8])
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sum(e.sim.launch_count for e in envs) - l0
    ms = ev0.elapsed_time(ev1)
    # ---------------- N > 1 only: the sharded layout's ONE collective — a single NCCL all-gather of the packed (obs, ref, reward,
    # terminated) buffer after every step (BASELINE.json north_star; SURVEY.md §8e asks for both figures) ----------------
    ms_gather, gather_bytes = None, 0
    if world > 1:
        from gym_electric_motor_b200.distributed import PackedStepOutputs

        packed = PackedStepOutputs(n, 14, 2, torch.float32, dev)
        gather_bytes = packed.nbytes
        gsim = envs[R - 1].sim
        gsim.bind_outputs(*packed.local_views())  # the kernel writes straight into the packed buffer
        for k in range(3):
            envs[R - 1].step(pool[k % 8])
            packed.gather_raw()
        barrier()
        eg0, eg1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eg0.record()
        for k in range(K):
            envs[R - 1].step(pool[k % 8])
            packed.gather_raw()
        eg1.record()
        barrier()
        ms_gather = eg0.elapsed_time(eg1)
    # ---------------- one launch at a time: events around every launch, L2 flushed before it (reference figure) ----------------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.zero_()
        evs[k][0].record()
        env.step(pool[k % 8])
        evs[k][1].record()
    barrier()
    ms_hot = sum(a.elapsed_time(b) for a, b in evs)
    # ---------------- e2e arm: host buffers through the C-ABI ----------------
    h_act = [torch.rand((n, 3), dtype=torch.float32).mul_(2).sub_(1).pin_memory() for _ in range(2)]
    h_obs = torch.empty((n, 14), dtype=torch.float32).pin_memory()
    h_ref = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_term = torch.empty(n, dtype=torch.uint8).pin_memory()
    ke = max(1, min(K, 50))
    for k in range(3):
        sim.step_host_ptr(h_act[k % 2].data_ptr(), h_obs.data_ptr(), h_ref.data_ptr(), h_rew.data_ptr(), h_term.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for k in range(ke):
        sim.step_host_ptr(h_act[k % 2].data_ptr(), h_obs.data_ptr(), h_ref.data_ptr(), h_rew.data_ptr(), h_term.data_ptr())
    barrier()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_hot, t_e2e * 1e3, ms_gather or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_hot, ms_e2e, ms_gather = [float(x) for x in t.tolist()]
    if rank == 0:
        total_envs = n * world
        ms_per_step = ms / K
        value = total_envs / (ms_per_step * 1e-3)
        peak, peak_src = peaks()
        achieved = B_ALG * n / (ms_per_step * 1e-3) / 1e9  # per GPU: this kernel's algorithmic GB/s
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "alg_bytes_per_env_step": B_ALG, "kernel": "step_kernel<SYNC,cont,f32,AoS>",
                         "kernel_ms": ms_per_step},
            "e2e": {"value": total_envs * ke / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * 3 * 4,
                    "d2h_bytes_per_step": n * (14 + 2 + 1) * 4 + n, "steps": ke, "ms_per_step": ms_e2e / ke,
                    "api": "gemb200_step_host via VectorSim.step_host_ptr (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cold_events": {"value": total_envs * K / (ms_hot * 1e-3), "ms_per_step": ms_hot / K,
                            "note": "one launch at a time, CUDA events around each launch, 256 MiB L2 flush before it"},
            "wall_ms_timed_region": t_wall * 1e3,
        }
        if world > 1:
            line["with_all_gather"] = {"value": total_envs * K / (ms_gather * 1e-3), "unit": UNIT, "ms_per_step": ms_gather / K,
                                       "bytes_per_rank_per_step": gather_bytes,
                                       "note": "every step followed by ONE NCCL all_gather_into_tensor of the packed (obs, ref, reward, terminated) "
                                               "buffer the kernel writes into; `value` above is the sharded layout without it (rank-local consumers)"}
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_arm(65536, 0, 2, budget_s=12.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "note": "python reference itself: ~8.6e3 env.step/s/core (BASELINE.md §2, survey container)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
