"""Host logic of ElectricMotorEnvironment around the device handle (scalar gym contract, callbacks, state filter, termination
assertion, reseeding, component setters), exercised on CPU with the handle class replaced by a scripted stand-in.  The stand-in does NO
physics — it returns scripted tensors — so nothing here is a parity statement; the real path is covered by the `-m gpu` tests."""
import numpy as np
import pytest
import torch

import gym_electric_motor_b200 as gem
from gym_electric_motor_b200 import _cabi as K


class ScriptedHandle:
    """the VectorSim surface core.py uses: reset(mask), step(action), close(), soa, set_reference"""

    created = []

    def __init__(self, cfg, reuse_outputs=True):
        self.cfg, self.n, self.soa, self.closed = cfg, cfg.n_envs, False, False
        dims = [K.C.c_int32() for _ in range(4)]
        K.check(K.load_library().gemb200_query_dims(K.C.byref(cfg), *[K.C.byref(d) for d in dims]), "query_dims")
        self.n_state, self.n_ref = dims[0].value, cfg.n_ref
        self.steps, self.terminate_at, self.actions = 0, None, []
        ScriptedHandle.created.append(self)

    def reset(self, mask=None):
        self.steps = 0
        return torch.arange(self.n * self.n_state, dtype=torch.float32).reshape(self.n, self.n_state) * 0.01, torch.full((self.n, self.n_ref), 0.5)

    def step(self, action):
        self.steps += 1
        self.actions.append(np.asarray(action.cpu() if torch.is_tensor(action) else action).copy())
        obs = torch.full((self.n, self.n_state), float(self.steps)) + torch.arange(self.n_state) * 0.01
        term = torch.zeros(self.n, dtype=torch.uint8)
        if self.terminate_at == self.steps:
            term[0] = 1
        return obs, torch.full((self.n, self.n_ref), 0.25), torch.full((self.n,), -0.5), term

    def reseed(self, seed):
        self.reseeds = getattr(self, "reseeds", []) + [int(seed)]
        self.cfg.seed = int(seed)

    def close(self):
        self.closed = True


@pytest.fixture
def scripted(monkeypatch):
    import gym_electric_motor_b200.vector_sim as vs

    ScriptedHandle.created = []
    monkeypatch.setattr(vs, "VectorSim", ScriptedHandle)
    return ScriptedHandle


class Recorder(gem.Callback):
    def __init__(self):
        self.events = []

    def set_env(self, env):
        self.events.append(("set_env", type(env).__name__))

    def on_reset_begin(self):
        self.events.append(("reset_begin",))

    def on_reset_end(self, state, reference):
        self.events.append(("reset_end", tuple(state.shape), tuple(reference.shape)))

    def on_step_begin(self, k, action):
        self.events.append(("step_begin", k))

    def on_step_end(self, k, state, reference, reward, terminated):
        self.events.append(("step_end", k))

    def on_close(self):
        self.events.append(("close",))


def test_scalar_contract_and_callbacks(scripted):
    rec = Recorder()
    env = gem.make("Cont-CC-PMSM-v0", callbacks=[rec], state_filter=["i_sd", "i_sq", "omega"])
    assert rec.events == [("set_env", "ContCurrentControlPermanentMagnetSynchronousMotorEnv")]
    with pytest.raises(AssertionError):  # reference core.py:341: a reset is required first
        env.step(np.zeros(3))
    (state, ref), info = env.reset()
    assert isinstance(state, np.ndarray) and state.dtype == np.float64 and state.shape == (3,) and ref.shape == (2,) and info == {}
    assert np.allclose(state, [0.05, 0.06, 0.0])                                  # the filter's order, not the system's
    handle = scripted.created[-1]
    handle.terminate_at = 2
    (state, ref), reward, terminated, truncated, info = env.step([0.1, 0.2, 0.3])
    assert isinstance(reward, float) and reward == -0.5 and terminated is False and truncated is False and info == {}
    assert handle.actions[-1].shape == (1, 3) and np.allclose(state, [1.05, 1.06, 1.0])
    assert env.physical_system.k == 1
    (_, _), _, terminated, _, _ = env.step([0, 0, 0])
    assert terminated is True
    with pytest.raises(AssertionError):  # stepping a terminated scalar env asserts like the reference
        env.step([0, 0, 0])
    env.reset()
    env.step([0, 0, 0])
    env.close()
    assert handle.closed
    kinds = [e[0] for e in rec.events]
    assert kinds == ["set_env", "reset_begin", "reset_end", "step_begin", "step_end", "step_begin", "step_end", "reset_begin", "reset_end",
                     "step_begin", "step_end", "close"]
    assert rec.events[2] == ("reset_end", (1, 3), (1, 2)) and rec.events[3] == ("step_begin", 0) and rec.events[4] == ("step_end", 1)


def test_batched_contract_reseed_and_setters(scripted):
    env = gem.make("Cont-SC-PermExDc-v0", num_envs=6, autoreset="same_step", seed=3)
    (state, ref), _ = env.reset()
    assert torch.is_tensor(state) and tuple(state.shape) == (6, 5) and tuple(ref.shape) == (6, 1)
    first = scripted.created[-1]
    assert first.cfg.seed == 3 and first.cfg.autoreset == K.AUTORESET_SAME_STEP and first.cfg.n_envs == 6
    first.terminate_at = 1
    (state, ref), reward, terminated, truncated, _ = env.step(torch.zeros(6, 1))
    assert terminated.dtype == torch.bool and terminated.tolist() == [True] + [False] * 5 and tuple(reward.shape) == (6,)
    env.step(torch.zeros(6, 1))                                      # batched envs with auto-reset keep stepping
    env.reset(seed=3)
    assert scripted.created[-1] is first and first.reseeds == [3]     # EVERY seeded reset re-keys the handle (reference core.py:300-304) ...
    env.reset(seed=3)
    assert first.reseeds == [3, 3]                                    # ... also with the seed it already has: equal seeds, identical episodes
    env.reset(seed=4)
    assert scripted.created[-1] is first and first.reseeds == [3, 3, 4] and first.cfg.seed == 4
    env.reset()
    assert first.reseeds == [3, 3, 4]                                 # an unseeded reset continues the streams
    second = scripted.created[-1]
    env.reference_generator = gem.reference_generators.ConstReferenceGenerator(reference_state="omega", reference_value=0.3)
    assert second.closed and env.physical_system._sim is None
    env.reset()
    third = scripted.created[-1]
    assert third is not second and third.cfg.ref_kind[0] == K.REF_CONST and third.cfg.ref_value[0] == 0.3
    env.close()
    assert third.closed


def test_vector_facade_flow(scripted):
    """gymnasium.vector conventions of gem.vector.make_vec: (obs, info) / 5-tuple, flattened Box observation, same-step auto-reset mode"""
    venv = gem.vector.make_vec("Cont-CC-PMSM-v0", num_envs=5, flatten_obs=True, seed=9)
    obs, info = venv.reset(seed=9)
    assert tuple(obs.shape) == (5, 16) and venv.observation_space.shape == (5, 16) and info == {}
    handle = scripted.created[-1]
    assert handle.cfg.autoreset == K.AUTORESET_SAME_STEP and venv.metadata["autoreset_mode"] == "same_step"
    handle.terminate_at = 1
    obs, rewards, terminations, truncations, infos = venv.step(torch.zeros(5, 3))
    assert tuple(obs.shape) == (5, 16) and terminations.tolist() == [True, False, False, False, False] and not truncations.any()
    assert torch.allclose(obs[:, 14:], torch.full((5, 2), 0.25)) and torch.allclose(rewards, torch.full((5,), -0.5))
    plain = gem.vector.make_vec("Finite-SC-PMSM-v0", num_envs=3)
    (state, ref), _ = plain.reset()
    assert tuple(state.shape) == (3, 14) and tuple(ref.shape) == (3, 1) and plain.single_action_space.n == 8
    venv.close(); plain.close()
    assert handle.closed and venv.closed


def test_scalar_env_asserts_invalid_finite_actions(scripted):
    """reference converters.py:204-206 / :827-829: a finite converter asserts that the action is an element of its action space"""
    env = gem.make("Finite-CC-PMSM-v0")
    env.reset()
    env.step(7)
    env.step(np.int64(0))
    for bad in (8, -1):
        with pytest.raises(AssertionError, match="not a valid element of the action space"):
            env.step(bad)
    multi = gem.make("Finite-CC-EESM-v0")  # MultiDiscrete([8, 4])
    multi.reset()
    multi.step([7, 3])
    with pytest.raises(AssertionError):
        multi.step([7, 4])
    cont = gem.make("Cont-CC-PMSM-v0")      # continuous converters clip instead (converters.py:160-163)
    cont.reset()
    cont.step([2.0, -3.0, 0.0])
