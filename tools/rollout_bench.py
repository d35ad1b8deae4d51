#!/usr/bin/env python
"""Device-side timing of the fused rollout (gemb200_rollout_record) against per-step launches, per configuration.

    python tools/rollout_bench.py [--config pmsm|fin_sc_pmsm|scim|eesm] [--out gpurun_out/rollout_bench.jsonl]

For every (N, K, record_every) the timed region is R launches back to back between one CUDA-event pair, rotating over replicas of the
env batch and over action / output tensors so that a launch never finds its data in L2 (working set per launch is printed).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def run(cfg_name, n, k, every, reps, out):
    spec = bench.CONFIGS[cfg_name]
    n_rep = max(2, min(4, (1 << 22) // n))
    envs = [bench.make_env(cfg_name, n, device=0, rank=r) for r in range(n_rep)]
    for e in envs:
        e.reset()
    sim0 = envs[0].sim
    dev = sim0.device
    n_act, n_obs, n_ref = sim0.n_act, sim0.n_state, sim0.n_ref
    gen = torch.Generator(device=dev).manual_seed(1)
    if sim0.finite:
        acts = [torch.randint(0, spec["n_finite"], (k, n, n_act), generator=gen, device=dev, dtype=torch.int32) for _ in range(2)]
    else:
        acts = [torch.rand((k, n, n_act), generator=gen, device=dev) * 2 - 1 for _ in range(2)]
    s = (k // every) if every else 1
    outs = [(torch.empty((s, n, n_obs), device=dev), torch.empty((s, n, max(n_ref, 1)), device=dev), torch.empty((s, n), device=dev),
             torch.empty((s, n), dtype=torch.uint8, device=dev)) for _ in range(2)]
    for r in range(3):
        envs[r % n_rep].sim.rollout_into(acts[r % 2], k, every, *outs[r % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        envs[r % n_rep].sim.rollout_into(acts[r % 2], k, every, *outs[r % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    b_io = n_act * 4 + (n_obs * 4 + n_ref * 4 + 5) * (1.0 / every if every else 1.0 / k)
    line = {"config": cfg_name, "n": n, "k": k, "record_every": every, "us_per_step": 1e3 * ms / k, "env_steps_per_s": n * k / (ms * 1e-3),
            "alg_bytes_per_env_step": b_io + spec["record_bytes"] / k, "GBps_alg": (b_io + spec["record_bytes"] / k) * n * k / (ms * 1e-3) / 1e9,
            "launch_ms": ms, "replicas": n_rep}
    print(json.dumps(line), flush=True)
    if out:
        with open(out, "a") as f:
            f.write(json.dumps(line) + "\n")
    for e in envs:
        e.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="pmsm")
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    grid = [(1 << 20, 1, 1), (1 << 20, 4, 1), (1 << 20, 16, 1), (1 << 20, 16, 0), (1 << 20, 16, 4), (1 << 20, 64, 0), (1 << 16, 1, 1), (1 << 16, 16, 1), (1 << 16, 64, 1),
            (1 << 16, 64, 0), (1 << 16, 256, 1)]
    if a.quick:
        grid = [(1 << 20, 16, 1), (1 << 20, 16, 0), (1 << 16, 64, 1)]
    for n, k, every in grid:
        run(a.config, n, k, every, max(4, 256 // k), a.out)


if __name__ == "__main__":
    main()
