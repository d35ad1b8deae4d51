#!/usr/bin/env python
"""How fast is the UNMODIFIED reference on the headline workload, next to the C oracle port that bench.py's cpu_baseline times?

Runs only in the build container (needs /root/reference; the GPU box has neither the reference nor gymnasium).  Same workload as
bench.py: Cont-CC-PMSM-v0, tau = 1e-4, random actions U(-1,1)^3, reset on termination.  Times env.step() of the reference with
(a) its default solver (scipy dopri5), (b) EulerSolver, (c) the test-side RK4 plugin used for the goldens — one process per core —
and the oracle port (RK4 x1) on the same cores.  One JSON line per arm -> profiles/r01_reference_cpu_here.jsonl
    python tools/time_reference_here.py [--steps 20000] [--procs N]"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_worker(args):
    solver, steps, seed = args
    sys.dont_write_bytecode = True
    warnings.filterwarnings("ignore")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as mg  # imports the reference through the test shims (no copy of reference code)
    import numpy as np

    gem = mg.gem
    kw = {}
    if solver == "euler":
        kw["ode_solver"] = mg.EulerSolver()
    elif solver == "rk4":
        kw["ode_solver"] = mg.RK4Solver(1)
    env = gem.make("Cont-CC-PMSM-v0", visualization=mg.NoViz(), **kw)
    env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, size=(steps, 3))
    for k in range(200):
        _, _, term, _, _ = env.step(acts[k])
        if term:
            env.reset()
    t0 = time.perf_counter()
    for k in range(steps):
        _, _, term, _, _ = env.step(acts[k])
        if term:
            env.reset()
    return steps / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = []
    for solver in ("dopri5", "euler", "rk4"):
        one = ref_worker((solver, a.steps, 0))
        with mp.get_context("spawn").Pool(a.procs) as pool:
            t0 = time.perf_counter()
            rates = pool.map(ref_worker, [(solver, a.steps, s) for s in range(a.procs)])
            wall = time.perf_counter() - t0
        out.append({"arm": f"reference env.step, {solver}", "steps_per_s_1core": one, "steps_per_s_all": sum(rates), "procs": a.procs,
                    "wall_s_incl_import": wall, "cpu": cpu, "where": "build container (not the GPU box)"})
        print(json.dumps(out[-1]), flush=True)
    sys.path.insert(0, ROOT)
    import bench

    v, cores, sample = bench.cpu_arm(65536, 0, 2, budget_s=10.0)
    out.append({"arm": "oracle port (C, fp64, RK4 x1), bench.py cpu_arm", "steps_per_s_all": v, "procs": cores, "sample": sample, "cpu": cpu,
                "where": "build container (not the GPU box)"})
    print(json.dumps(out[-1]), flush=True)
    with open(os.path.join(ROOT, "profiles", "r01_reference_cpu_here.jsonl"), "w") as f:
        for o in out:
            f.write(json.dumps(o) + "\n")


if __name__ == "__main__":
    main()
