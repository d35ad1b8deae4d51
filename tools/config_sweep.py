"""Throughput of the BASELINE.json configs beyond the headline one (kernel time via CUDA events, L2 flushed per step).
Writes one JSON line per config to stdout.   python tools/config_sweep.py [--n 1048576] [--steps 100]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gym_electric_motor_b200 as gem  # noqa: E402
from gym_electric_motor_b200.mixed import MixedEnvBatch  # noqa: E402

B_ALG = {"PMSM": 129, "SynRM": 129, "SCIM": 145, "EESM": 161, "FinPMSM": 109}
PEAK = 6568.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def actions_for(env, n, dev, gen):
    sp = env.action_space
    if hasattr(sp, "nvec"):
        return torch.stack([torch.randint(0, int(k), (n,), generator=gen, device=dev, dtype=torch.int32) for k in sp.nvec], dim=1).contiguous()
    if hasattr(sp, "n"):
        return torch.randint(0, sp.n, (n, 1), generator=gen, device=dev, dtype=torch.int32)
    dt = torch.float64 if env.sim.dtype == torch.float64 else torch.float32
    lo = torch.as_tensor(sp.low, device=dev, dtype=dt)
    hi = torch.as_tensor(sp.high, device=dev, dtype=dt)
    return (torch.rand((n, len(sp.low)), generator=gen, device=dev, dtype=dt) * (hi - lo) + lo).contiguous()


def time_steps(step_fn, steps, flush):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(10):
        flush.zero_()
        step_fn(k)
    torch.cuda.synchronize()
    for k in range(steps):
        flush.zero_()
        evs[k][0].record()
        step_fn(k)
        evs[k][1].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n = args.n
    RK4 = gem.physical_systems.RK4Solver
    cases = [
        ("Cont-CC-PMSM-v0 rk4x1 f32", "Cont-CC-PMSM-v0", dict(ode_solver=RK4()), 129),
        ("Cont-CC-PMSM-v0 rk4x2 (default scipy mapping) f32", "Cont-CC-PMSM-v0", dict(), 129),
        ("Cont-CC-PMSM-v0 euler f32", "Cont-CC-PMSM-v0", dict(ode_solver=gem.physical_systems.EulerSolver()), 129),
        ("Cont-CC-PMSM-v0 rk4x1 f32 SoA obs", "Cont-CC-PMSM-v0", dict(ode_solver=RK4(), layout="soa"), 129),
        ("Cont-CC-PMSM-v0 rk4x1 f64", "Cont-CC-PMSM-v0", dict(ode_solver=RK4(), dtype="float64"), 129 * 2),
        ("Finite-SC-PMSM-v0 rk4x1 f32", "Finite-SC-PMSM-v0", dict(ode_solver=RK4()), 109),
        ("Finite-SC-PMSM-v0 rk4x1 interlock 1us f32", "Finite-SC-PMSM-v0", dict(ode_solver=RK4(), converter=dict(interlocking_time=1e-6)), 109),
        ("Cont-CC-SCIM-v0 rk4x1 f32", "Cont-CC-SCIM-v0", dict(ode_solver=RK4()), 145),
        ("Cont-CC-EESM-v0 rk4x1 f32", "Cont-CC-EESM-v0", dict(ode_solver=RK4()), 161),
        ("Cont-CC-SynRM-v0 rk4x1 f32", "Cont-CC-SynRM-v0", dict(ode_solver=RK4()), 129),
        ("Cont-SC-PMSM-v0 rk4x1 f32 (PolynomialStaticLoad)", "Cont-SC-PMSM-v0", dict(ode_solver=RK4()), 117),
        ("Cont-CC-DFIM-v0 rk4x1 f32", "Cont-CC-DFIM-v0", dict(ode_solver=RK4()), 4 * 6 + 2 * 4 * 6 + 4 * 24 + 4 * 2 + 2 * 4 * 2 + 5),
        ("Cont-CC-PermExDc-v0 euler f32", "Cont-CC-PermExDc-v0", dict(ode_solver=gem.physical_systems.EulerSolver()), 4 + 2 * 4 * 2 + 4 * 5 + 4 + 2 * 4 + 5),
    ]
    for label, env_id, kw, balg in cases:
        env = gem.make(env_id, num_envs=n, autoreset="same_step", seed=0, **kw)
        env.reset()
        pool = [actions_for(env, n, dev, gen) for _ in range(4)]
        if env.sim.soa:
            pool = [a.T.contiguous() for a in pool]
        ms = time_steps(lambda k: env.step(pool[k % 4]), args.steps, flush)
        print(json.dumps({"config": label, "n_envs": n, "ms_per_step": ms, "env_steps_per_s": n / (ms * 1e-3), "alg_bytes_per_env_step": balg,
                          "alg_gbs": balg * n / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": balg * n / (ms * 1e-3) / 1e9 / PEAK}), flush=True)
        env.close()
    # BASELINE.json configs[1]: Cont-CC-PMSM-v0 at N = 65536 (one partial wave: latency- and launch-bound, not bandwidth-bound):
    # K back-to-back launches issued from C (gemb200_rollout), one event pair; RK4 x1 and x2
    for label, kw in (("Cont-CC-PMSM-v0 N=65536 rk4x1 f32 (64 back-to-back launches)", dict(ode_solver=RK4())),
                      ("Cont-CC-PMSM-v0 N=65536 rk4x2 f32 (64 back-to-back launches)", dict(ode_solver=RK4(nsteps=2)))):
        ns = 1 << 16
        env = gem.make("Cont-CC-PMSM-v0", num_envs=ns, autoreset="same_step", seed=0, **kw)
        env.reset()
        roll = torch.stack([actions_for(env, ns, dev, gen) for _ in range(64)]).contiguous()
        for _ in range(3):
            env.sim.rollout(roll)
        best = 1e9
        for _ in range(5):
            env.sim.time_begin(); env.sim.rollout(roll); best = min(best, env.sim.time_end() / 64)
        print(json.dumps({"config": label, "n_envs": ns, "ms_per_step": best, "env_steps_per_s": ns / (best * 1e-3), "alg_bytes_per_env_step": 129,
                          "alg_gbs": 129 * ns / (best * 1e-3) / 1e9, "frac_of_measured_hbm": 129 * ns / (best * 1e-3) / 1e9 / PEAK}), flush=True)
        env.close()
    # configs[4]: mixed PMSM + SynRM + EESM, interleaved ids, per-type kernels on separate streams
    ids = [("Cont-CC-PMSM-v0", dict(ode_solver=RK4())), ("Cont-CC-SynRM-v0", dict(ode_solver=RK4())), ("Cont-CC-EESM-v0", dict(ode_solver=RK4()))]
    nm = (n // 3) * 3
    mixed = MixedEnvBatch(ids, nm, autoreset="same_step", seed=0)
    mixed.reset()
    pools = [[actions_for(e, mixed.per_type, dev, gen) for e in mixed.envs] for _ in range(4)]
    ms = time_steps(lambda k: mixed.step(pools[k % 4]), args.steps, flush)
    balg = (129 + 129 + 161) / 3
    print(json.dumps({"config": "mixed 1/3 PMSM + 1/3 SynRM + 1/3 EESM, 3 streams", "n_envs": nm, "ms_per_step": ms, "env_steps_per_s": nm / (ms * 1e-3),
                      "alg_bytes_per_env_step": balg, "alg_gbs": balg * nm / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": balg * nm / (ms * 1e-3) / 1e9 / PEAK}), flush=True)


if __name__ == "__main__":
    main()
