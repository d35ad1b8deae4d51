"""A few steps of every feature path at a ragged batch size — meant to run under `compute-sanitizer --tool memcheck` on the GPU box:
    compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gym_electric_motor_b200 as gem  # noqa: E402

ps, psw, rg = gem.physical_systems, gem.physical_system_wrappers, gem.reference_generators
RK4, N = ps.RK4Solver, 777


def profile(t, a, f, o):
    return o + a * np.sin(2 * np.pi * f * t)


CASES = [
    ("Cont-CC-PMSM-v0", dict(ode_solver=RK4())),                                                    # PLAIN
    ("Cont-CC-PMSM-v0", dict(ode_solver=RK4(), layout="soa")),
    ("Cont-CC-PMSM-v0", dict(ode_solver=RK4(), dtype="float64")),
    ("Finite-SC-PMSM-v0", dict(ode_solver=RK4(), converter=dict(interlocking_time=1e-6))),
    ("Cont-CC-EESM-v0", dict(ode_solver=ps.EulerSolver(nsteps=3), physical_system_wrappers=[psw.DeadTimeProcessor(steps=2), psw.DqToAbcActionProcessor.make("EESM")])),
    ("Cont-SC-SCIM-v0", dict(ode_solver=RK4(), physical_system_wrappers=[psw.FluxObserver(), psw.CosSinProcessor(angle="psi_angle"),
                                                                           psw.StateNoiseProcessor(states="all", random_kwargs=dict(scale=1e-3)),
                                                                           psw.DqToAbcActionProcessor.make("SCIM")])),
    ("Cont-SC-PMSM-v0", dict(ode_solver=RK4(), physical_system_wrappers=[psw.CosSinProcessor(remove_angle=True)], layout="soa")),
    ("Cont-CC-DFIM-v0", dict(ode_solver=RK4(), physical_system_wrappers=[psw.FluxObserver(), psw.DqToAbcActionProcessor.make("DFIM")])),
    ("Finite-CC-DFIM-v0", dict(ode_solver=RK4())),
    ("Cont-SC-PermExDc-v0", dict(ode_solver=RK4(), supply=ps.RCVoltageSupply(60.0, dict(R=0.5, C=4e-3)))),
    ("Finite-CC-ExtExDc-v0", dict(ode_solver=RK4(), supply=ps.RCVoltageSupply(60.0, dict(R=0.5, C=4e-3)))),
    ("Cont-CC-ShuntDc-v0", dict(ode_solver=RK4(), supply=ps.AC1PhaseSupply(42.0, dict(frequency=50.0)))),
    ("Cont-CC-PMSM-v0", dict(ode_solver=RK4(nsteps=2), load=ps.ExternalSpeedLoad(profile, speed_profile_kwargs=dict(a=50.0, f=20.0, o=100.0), horizon_steps=64))),
    ("Cont-SC-SynRM-v0", dict(ode_solver=RK4(), motor=dict(motor_initializer=dict(random_init="gaussian")), load=dict(load_initializer=dict(random_init="uniform")),
                              reference_generator=rg.SwitchedReferenceGenerator([rg.WienerProcessReferenceGenerator(reference_state="omega"),
                                                                                 rg.SinusoidalReferenceGenerator(reference_state="omega"),
                                                                                 rg.StepReferenceGenerator(reference_state="omega")], super_episode_length=(2, 5)))),
    ("Cont-TC-SeriesDc-v0", dict(ode_solver=RK4(), reference_generator=rg.LaplaceProcessReferenceGenerator(reference_state="torque"))),
    # round 2
    ("Cont-SC-SCIM-v0", dict(ode_solver=RK4(), motor=dict(motor_initializer=dict(random_init="uniform")), load=dict(load_initializer=dict(random_init="uniform", interval=[[-50.0, 120.0]])))),
    ("Finite-CC-ExtExDc-v0", dict(ode_solver=RK4(), converter=ps.FiniteMultiConverter([ps.FiniteFourQuadrantConverter(interlocking_time=1e-6),
                                                                                         ps.FiniteFourQuadrantConverter(interlocking_time=2.5e-6)]))),
    ("Cont-CC-ExtExDc-v0", dict(ode_solver=RK4(), physical_system_wrappers=[psw.CurrentSumProcessor(currents=["i_a", "i_e"], limit="sum")])),
    ("Cont-SC-PMSM-v0", dict(ode_solver=RK4(), _env_params=True)),
]
if os.environ.get("SANITIZE_CASES"):  # e.g. "0,1,5,7" for the slower racecheck tool
    CASES = [CASES[int(k)] for k in os.environ["SANITIZE_CASES"].split(",")]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
for env_id, kw in CASES:
    kw = dict(kw)
    env_params = kw.pop("_env_params", False)
    env = gem.make(env_id, num_envs=N, autoreset="same_step", seed=1, **kw)
    if env_params:  # per-env parameter blocks (ENVP instantiation)
        env.set_env_parameters(motor_parameter=dict(r_s=np.linspace(15e-3, 25e-3, N), psi_p=np.linspace(0.15, 0.19, N)), load_parameter=dict(j_load=np.linspace(1e-4, 3e-4, N)))
    env.reset()
    sp = env.action_space
    for k in range(6):
        if hasattr(sp, "nvec"):
            a = torch.stack([torch.randint(0, int(m), (N,), generator=gen, device=dev, dtype=torch.int32) for m in sp.nvec], dim=1).contiguous()
        elif hasattr(sp, "n"):
            a = torch.randint(0, sp.n, (N, 1), generator=gen, device=dev, dtype=torch.int32)
        else:
            dt = torch.float64 if env.sim.dtype == torch.float64 else torch.float32
            a = torch.rand((N, len(sp.low)), generator=gen, device=dev, dtype=dt) * 2 - 1
        if env.sim.soa:
            a = a.T.contiguous()
        (obs, ref), rew, term, _, _ = env.step(a)
    if not env.sim.soa:  # fused rollout: 5 steps in one launch, every step recorded, then 3 with only the last one
        acts = torch.stack([a] * 5).contiguous()
        (st, rf), rw_, tm = env.rollout(acts, record_every=1)
        env.rollout(acts[:3], record_every=0)
        assert torch.isfinite(st).all()
        # round 2, second half: a subset of the outputs (the kernels' per-output store path), then the device-resident clock (clock read from
        # device memory, one-thread tick kernel) for steps, a rollout and a masked reset, and back to the host clock
        sim = env.sim
        part = [torch.empty_like(st), None, torch.empty_like(rw_), None]
        sim.rollout_into(acts, 5, 1, *part)
        sim.set_device_clock(True)
        env.step(a)
        env.rollout(acts[:4], record_every=2)
        env.reset(mask=torch.ones(N, dtype=torch.uint8, device=dev))
        env.step(a)
        sim.set_device_clock(False)
    env.reset(seed=3)  # gemb200_reseed
    env.reset(mask=torch.ones(N, dtype=torch.uint8, device=dev))
    sd = env.state_dict()
    env.load_state_dict(sd)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all(), env_id
    print("ok", env_id, list(obs.shape), flush=True)
    env.close()
# the multi-destination store path (gemb200_bind_peers) with one rank: outputs go through the destination list into a library-allocated buffer
from gym_electric_motor_b200.distributed import PeerGather  # noqa: E402

env = gem.make("Cont-CC-PMSM-v0", num_envs=N, autoreset="same_step", seed=1, ode_solver=RK4())
env.reset()
pg = PeerGather(env.sim, torch.float32)
for k in range(4):
    b = pg.step(torch.rand((N, 3), generator=gen, device=dev) * 2 - 1)
pg.finish()
pg.check()
assert torch.isfinite(pg.views(b)[0]).all()
pg.release()
env.close()
print("ok peer stores (1 rank)")
print("all paths ran")
