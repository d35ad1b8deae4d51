"""Steady-state DRAM traffic of the step / rollout kernels, for `ncu --replay-mode range`:

    ncu --replay-mode range --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum \
        --csv --log-file gpurun_out/traffic_<mode>_<config>.csv python tools/traffic_range.py <mode> <config> [launches]

mode = step     : `launches` single-step launches rotating over 4 replicas of the env batch (every launch finds its records in HBM);
mode = rollout  : `launches` fused launches of 16 steps, every step recorded.
The profiled range holds ONLY those launches (cudaProfilerStart/Stop); the same number of identical launches runs right before it, so the
write-back traffic that is still in L2 when the range ends is balanced by what the launches before the range left behind: the counters
are the steady-state bytes.  Prints the env-steps inside the range so that bytes per env-step can be formed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "step"
name = sys.argv[2] if len(sys.argv) > 2 else "pmsm"
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n = bench.CONFIGS[name]["envs"]
K = 16
if mode == "step":
    R = 4
    envs = [bench.make_env(name, n, device=0, rank=r) for r in range(R)]
    for e in envs:
        e.reset()
    sim = envs[0].sim
    gen = torch.Generator(device="cuda").manual_seed(0)
    if sim.finite:
        pool = [torch.randint(0, 8, (n, sim.n_act), generator=gen, device="cuda", dtype=torch.int32) for _ in range(8)]
    else:
        pool = [torch.rand((n, sim.n_act), generator=gen, device="cuda") * 2 - 1 for _ in range(8)]

    def go(cnt):
        for k in range(cnt):
            envs[k % R].step(pool[k % 8])
    steps = launches
else:
    wl = bench.Workload(name, n, 0, 0, torch)

    def go(cnt):
        for _ in range(cnt):
            wl.rollout(K, 1)
    steps = launches * K
go(max(launches, 4))
torch.cuda.synchronize()
torch.cuda.profiler.start()
go(launches)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"RANGE mode={mode} config={name} envs={n} launches={launches} env_batch_steps={steps}")
