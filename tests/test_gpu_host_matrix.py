"""Kernel vs oracle for every configuration of the user-kwargs matrix (tests/agent_surface/kwargs_matrix_harness.py: 95 kwarg combinations +
the 54 default ids), built through the HOST API (`gem.make(**kwargs)` -> `env.build_config()`), float64 build, 6 envs, 12 steps with fixed
actions: states, rewards, terminations, next references.

Marked `gpu_next`, NOT `gpu`: these combinations were assembled after the round's GPU budget was spent, so they have never run on a
device.  The first thing to do with a GPU is `pytest -m gpu_next`; cases that pass move under the `gpu` marker.  (The reference-vs-oracle
half of the same matrix runs on CPU in tests/test_agent_surface.py.)"""
import numpy as np
import pytest

from gym_electric_motor_b200 import _cabi as K

pytestmark = pytest.mark.gpu_next


def _cases():
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "agent_surface"))
    import kwargs_matrix_harness as h

    return h.CASES, h.PRELUDE


CASES, PRELUDE = _cases()
REFUSED = ("interlock_cont_multi", "finite_multi_interlock", "currentsum_extex", "synrm_dq")  # refused on the host (tests/test_agent_surface.py)


@pytest.mark.parametrize("case", sorted(c for c in CASES if c not in REFUSED and not c.startswith("err_")))
def test_device_matches_oracle_for_host_built_config(oracle_lib, case):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (`-m gpu_next` on the GPU box)")
    import gym_electric_motor_b200 as gem
    from gym_electric_motor_b200.vector_sim import VectorSim

    ns = {"gem": gem}
    exec(PRELUDE, ns)
    env = eval(CASES[case], ns)
    cfg = env.build_config()
    n = 6
    cfg.n_envs, cfg.dtype, cfg.seed, cfg.autoreset = n, K.F64, 5, K.AUTORESET_SAME_STEP
    ora, sim = oracle_lib.Oracle(cfg), VectorSim(cfg)
    o_obs, o_ref = ora.reset()
    d_obs, d_ref = sim.reset()
    assert np.abs(d_obs.double().cpu().numpy() - o_obs).max() < 1e-9 and np.abs(d_ref.double().cpu().numpy() - o_ref).max() < 1e-9
    rng = np.random.default_rng(3)
    sp = env.action_space
    for k in range(12):
        if hasattr(sp, "low"):
            a = rng.uniform(sp.low, sp.high, size=(n, len(sp.low)))
        elif hasattr(sp, "nvec"):
            a = np.stack([rng.integers(0, int(m), size=n) for m in sp.nvec], axis=1).astype(np.int32)
        else:
            a = rng.integers(0, sp.n, size=(n, 1)).astype(np.int32)
        o = ora.step(a)
        d = sim.step(a)
        tol = 1e-8
        assert np.abs(d[0].double().cpu().numpy() - o[0]).max() < tol, (case, k)
        assert np.abs(d[1].double().cpu().numpy() - o[1]).max() < tol, (case, k)
        assert np.abs(d[2].double().cpu().numpy() - o[2]).max() < 10 * tol, (case, k)
        assert np.array_equal(d[3].cpu().numpy().astype(np.uint8), o[3]), (case, k)
    sim.close()
