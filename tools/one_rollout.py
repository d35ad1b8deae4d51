"""One fused rollout launch (and a few single steps) of a bench config — the target of `ncu` captures.
usage: python tools/one_rollout.py [config] [K] [record_every] [n_envs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pmsm"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
every = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = int(sys.argv[4]) if len(sys.argv) > 4 else bench.CONFIGS[name]["envs"]
bench.ROLL_MAX = max(bench.ROLL_MAX, k)
wl = bench.Workload(name, n, 0, 0, torch)
for _ in range(2):
    wl.rollout(k, every)
torch.cuda.synchronize()
if not wl.mixed:
    for j in range(4):
        wl.env.step(wl.acts[0][j])
torch.cuda.synchronize()
print("done", name, n, k, every)
