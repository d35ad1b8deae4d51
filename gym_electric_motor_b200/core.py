"""Batched `ElectricMotorEnvironment` — the drop-in for reference core.py:53-392.

`reset()` / `step()` keep the reference's return structure
    reset -> ((state, reference_observation), info)
    step  -> ((state, reference_observation), reward, terminated, truncated, info)
with a leading env dimension N on every array (torch tensors on the CUDA device).  With `num_envs=None` the
environment runs ONE env and converts to the reference's scalar contract (numpy vectors, float reward, bool
terminated, assertion on stepping a terminated env, core.py:341).

The whole step — physical system, reference lookup, constraint check, reward, next reference — is one CUDA launch
through the C-ABI (include/gemb200.h: gemb200_step); see DESIGN.md.
"""
import numpy as np
import torch

from . import _cabi as K
from .constraints import Constraint, ConstraintMonitor
from .reference_generators import ReferenceGenerator
from .reward_functions import RewardFunction
from .spaces import Box, Tuple

try:  # a gymnasium.Env when gymnasium is installed, so that gymnasium.Wrapper / TimeLimit / FlattenObservation accept the scalar env
    from gymnasium import Env as _EnvBase  # pragma: no cover - this image has no gymnasium
except ImportError:
    _EnvBase = object


class Callback:
    """reference core.py:708-740 (hooks receive batched tensors)"""

    _env = None

    def set_env(self, env):
        self._env = env

    def on_reset_begin(self):
        pass

    def on_reset_end(self, state, reference):
        pass

    def on_step_begin(self, k, action):
        pass

    def on_step_end(self, k, state, reference, reward, terminated):
        pass

    def on_close(self):
        pass


class ElectricMotorVisualization(Callback):
    """reference core.py:743-753.  Plotting is host-side and out of scope; subclasses may still hook in."""

    def render(self):
        pass


class ElectricMotorEnvironment(_EnvBase):
    """See module docstring.  Constructor signature follows reference core.py:197-209 plus the batch options
    `num_envs`, `autoreset` ('same_step' | None), `seed`."""

    metadata = {}
    render_mode = None
    env_id = None

    def __init__(self, physical_system, reference_generator, reward_function, visualization=(), state_filter=None, callbacks=(),
                 constraints=(), physical_system_wrappers=(), scale_plots=False, num_envs=None, autoreset=None, seed=None, **kwargs):
        physical_system.apply_wrappers(tuple(physical_system_wrappers))  # fused into the kernel (physical_system_wrappers.py)
        if not isinstance(reference_generator, ReferenceGenerator):
            raise TypeError("reference_generator must be a built-in gym_electric_motor_b200 ReferenceGenerator")
        if not isinstance(reward_function, RewardFunction):
            raise TypeError("reward_function must be a built-in gym_electric_motor_b200 RewardFunction")
        self._scalar = num_envs is None
        if self._scalar and getattr(physical_system, "_layout", K.LAYOUT_AOS) == K.LAYOUT_SOA:
            raise ValueError("layout='soa' needs a batched environment (num_envs=...): the scalar contract returns one row per step")
        self._physical_system = physical_system
        self._reference_generator = reference_generator
        self._reward_function = reward_function
        self.num_envs = physical_system.num_envs
        if isinstance(constraints, ConstraintMonitor):
            cm = constraints
        else:
            limit_constraints = [c for c in constraints if isinstance(c, str)]
            additional = [c for c in constraints if isinstance(c, Constraint)]
            bad = [c for c in constraints if not isinstance(c, (str, Constraint))]
            if bad:
                raise TypeError("callable constraints are host code and cannot be fused into the kernel epilogue")
            cm = ConstraintMonitor(limit_constraints, additional)
        self._constraint_monitor = cm
        self._reference_generator.set_modules(self._physical_system)
        self._constraint_monitor.set_modules(self._physical_system)
        self._reward_function.set_modules(self._physical_system, self._reference_generator, self._constraint_monitor)
        ps = self._physical_system
        state_filter = state_filter or ps.state_names
        self.state_filter = [ps.state_names.index(s) for s in state_filter]
        self._filter_identity = self.state_filter == list(range(len(ps.state_names)))
        state_space = Box(ps.state_space.low[self.state_filter], ps.state_space.high[self.state_filter], dtype=np.float64)
        self.observation_space = Tuple((state_space, self._reference_generator.reference_space))
        self.action_space = ps.action_space
        self.reward_range = self._reward_function.reward_range
        self._terminated = True
        self._truncated = False
        self.scale_plots = scale_plots
        if isinstance(visualization, ElectricMotorVisualization):
            visualization = [visualization]
        self._visualizations = [v for v in (visualization or []) if isinstance(v, ElectricMotorVisualization)]
        self._callbacks = list(callbacks) + list(self._visualizations)
        if not self._visualizations:  # the reference's envs default to a MotorDashboard (user code reads env.visualizations[0]); here an
            from .visualization import MotorDashboard  # inert one that is not even registered as a callback (nothing to call per step)

            self._visualizations = [MotorDashboard(_quiet=True)]
        self._autoreset = K.AUTORESET_SAME_STEP if (autoreset in ("same_step", True, K.AUTORESET_SAME_STEP)) else K.AUTORESET_NONE
        self._seed_value = 0 if seed is None else int(seed)
        self._sim = None
        self._filter_index = None
        self._call_callbacks("set_env", self)

    # ------------------------------------------------------------------ reference-compatible properties
    @property
    def physical_system(self):
        return self._physical_system

    @property
    def reference_generator(self):
        return self._reference_generator

    @reference_generator.setter
    def reference_generator(self, reference_generator):
        """reference core.py:132-142: a new generator, then a reset is required.  Here the generator is part of the device handle's
        configuration, so the handle is dropped and rebuilt at the next reset."""
        if not isinstance(reference_generator, ReferenceGenerator):
            raise TypeError("reference_generator must be a built-in gym_electric_motor_b200 ReferenceGenerator")
        self._reference_generator = reference_generator
        self._reference_generator.set_modules(self._physical_system)
        self._reward_function.set_modules(self._physical_system, self._reference_generator, self._constraint_monitor)
        self.observation_space = Tuple((self.observation_space.spaces[0], self._reference_generator.reference_space))
        self._drop_handle()

    @property
    def reward_function(self):
        return self._reward_function

    @reward_function.setter
    def reward_function(self, reward_function):
        """reference core.py:153-162"""
        if not isinstance(reward_function, RewardFunction):
            raise TypeError("reward_function must be a built-in gym_electric_motor_b200 RewardFunction")
        self._reward_function = reward_function
        self._reward_function.set_modules(self._physical_system, self._reference_generator, self._constraint_monitor)
        self.reward_range = self._reward_function.reward_range
        self._drop_handle()

    def _drop_handle(self):
        self._terminated = True
        if self._sim is not None:
            self._sim.close()
            self._sim = None
            self._physical_system.attach(None)

    @property
    def constraint_monitor(self):
        return self._constraint_monitor

    @property
    def visualizations(self):
        """reference core.py:140-143 (agents look for a MotorDashboard in this list)"""
        return self._visualizations

    @property
    def limits(self):
        """limits of the states the env returns, i.e. after the state filter (reference core.py:169-174)"""
        return self._physical_system.limits[self.state_filter]

    @property
    def state_names(self):
        return [self._physical_system.state_names[s] for s in self.state_filter]

    @property
    def reference_names(self):
        return self._reference_generator.reference_names

    @property
    def nominal_state(self):
        return self._physical_system.nominal_state[self.state_filter]

    @property
    def unwrapped(self):
        return self

    @property
    def sim(self):
        """The underlying device handle wrapper (VectorSim)."""
        return self._ensure_sim()

    # ------------------------------------------------------------------ device handle
    def build_config(self):
        cfg = self._physical_system.fill_config(K.new_config())
        self._constraint_monitor.fill_config(cfg)
        self._reward_function.fill_config(cfg)
        self._reference_generator.fill_config(cfg)
        cfg.autoreset = self._autoreset
        cfg.seed = self._seed_value & 0xFFFFFFFFFFFFFFFF
        return cfg

    def _ensure_sim(self):
        if self._sim is None:
            from .vector_sim import VectorSim

            self._sim = VectorSim(self.build_config())
            self._physical_system.attach(self._sim)
        return self._sim

    def _call_callbacks(self, func_name, *args):
        for callback in self._callbacks:
            getattr(callback, func_name)(*args)

    def _filter(self, obs):
        if self._filter_identity:
            return obs
        import torch

        if self._filter_index is None:
            self._filter_index = torch.as_tensor(self.state_filter, device=obs.device)
        dim = 0 if self._sim.soa else 1
        return obs.index_select(dim, self._filter_index)

    # ------------------------------------------------------------------ gym API
    def reset(self, seed=None, options=None, mask=None, *_, **__):
        """core.py:300-319.  `seed` re-keys the device RNG streams (new handle state); `mask` (batched mode only)
        resets a subset of envs and returns the full observation tensors."""
        reseed = seed is not None and self._sim is not None
        if seed is not None:
            self._seed_value = int(seed)
        sim = self._ensure_sim()
        if reseed:  # the reference re-seeds every component on EVERY seeded reset (core.py:300-304): equal seeds, identical episodes
            sim.reseed(self._seed_value)
        self._call_callbacks("on_reset_begin")
        obs, ref = sim.reset(mask)
        self._terminated = False
        self._physical_system._k = 0
        state = self._filter(obs)
        self._call_callbacks("on_reset_end", state, ref)
        if self._scalar:
            return (state.double().cpu().numpy()[0], ref.double().cpu().numpy()[0]), {}
        return (state, ref), {}

    def step(self, action):
        """core.py:328-371"""
        sim = self._ensure_sim()
        if self._scalar:
            assert not self._terminated, "A reset is required before the environment can perform further steps"
            if not hasattr(self.action_space, "low"):  # finite converters assert their action (converters.py:204-206, :356-358, :827-829)
                candidate = np.asarray(action)
                candidate = int(candidate.reshape(-1)[0]) if candidate.size == 1 and hasattr(self.action_space, "n") else candidate.reshape(-1)
                assert self.action_space.contains(candidate), \
                    f"The selected action {action} is not a valid element of the action space {self.action_space}."
            action = np.asarray(action).reshape(1, -1)
        if self._callbacks:
            self._call_callbacks("on_step_begin", self._physical_system.k, action)
        obs, ref, reward, terminated = sim.step(action)
        self._physical_system._k += 1
        state = obs if self._filter_identity else self._filter(obs)
        if self._callbacks:
            self._call_callbacks("on_step_end", self._physical_system.k, state, ref, reward, terminated)
        if self._scalar:
            term = bool(terminated[0].item())
            self._terminated = term and self._autoreset == K.AUTORESET_NONE
            return (state.double().cpu().numpy()[0], ref.double().cpu().numpy()[0]), float(reward[0].item()), term, self._truncated, {}
        return (state, ref), reward, terminated.view(torch.bool), self._truncated, {}  # uint8 0/1 reinterpreted, no kernel

    def rollout(self, actions, record_every=1):
        """K consecutive `step` calls with pre-computed actions [K, N, n_act] in ONE kernel launch (open loop; bit-identical to calling
        `step` K times, core.py:328-371).  Returns ((states, references), rewards, terminateds) with a leading axis of K // record_every
        recorded steps (record_every = 0: only the last step, without the leading axis).  Batched mode only; callbacks see no
        per-step hooks."""
        if self._scalar:
            raise TypeError("rollout() needs a batched environment (num_envs=...)")
        sim = self._ensure_sim()
        obs, ref, reward, terminated = sim.rollout(actions, record_every)
        k = int(actions.shape[0]) if hasattr(actions, "shape") else len(actions)
        self._physical_system._k += k
        if not self._filter_identity:
            if self._filter_index is None:
                self._filter_index = torch.as_tensor(self.state_filter, device=obs.device)
            dim = (0 if sim.soa else 1) + (1 if record_every else 0)
            obs = obs.index_select(dim, self._filter_index)
        return (obs, ref), reward, terminated.view(torch.bool)

    def capture_steps(self, policy, n_steps, record=False, warmup=1):
        """`n_steps` closed-loop steps — action = policy(state, reference); env.step(action) — captured ONCE in a CUDA graph (graph.py);
        `.replay()` of the returned object runs them with a single call.  Batched mode only."""
        from .graph import CapturedSteps

        return CapturedSteps(self, policy, n_steps, record=record, warmup=warmup)

    _MP_SLOT = dict(p=K.MP_P, r_s=K.MP_R_S, l_d=K.MP_L_D, l_q=K.MP_L_Q, psi_p=K.MP_PSI_P, j_rotor=K.MP_J_ROTOR, r_a=K.MP_R_A, l_a=K.MP_L_A, psi_e=K.MP_PSI_E,
                    r_e=K.MP_R_E, l_e=K.MP_L_E, l_e_prime=K.MP_L_E_PRIME, l_m=K.MP_L_M, k=K.MP_K, l_sigs=K.MP_L_SIGS, l_sigr=K.MP_L_SIGR, r_r=K.MP_R_E)
    _LP_SLOT = dict(a=K.LP_A, b=K.LP_B, c=K.LP_C, j_load=K.LP_J_LOAD)

    def set_env_parameters(self, motor_parameter=None, load_parameter=None):
        """Domain randomisation: give every env of the batch its own physical parameters — the batched counterpart of constructing N
        reference envs with N `motor_parameter` / `load_parameter` dicts (electric_motor.py:118-131, polynomial_static_load.py:46-64).
        Both arguments are dicts  name -> array of N values  with the reference's parameter names (`r_s`, `l_d`, `psi_p`, `j_rotor`, ...;
        `a`, `b`, `c`, `j_load`); parameters that are not named keep the value the env was made with.  Limits, nominal values and the
        normalisation stay those of the env.  `set_env_parameters()` without arguments returns to the shared parameters."""
        if self._scalar:
            raise TypeError("set_env_parameters() needs a batched environment (num_envs=...)")
        sim = self._ensure_sim()
        if not motor_parameter and not load_parameter:
            sim.set_env_params(None, None)
            return
        cfg = sim.cfg
        mp = np.tile(np.array(list(cfg.motor_param), dtype=np.float64), (sim.n, 1))
        lp = np.tile(np.array(list(cfg.load_param), dtype=np.float64), (sim.n, 1))
        for name, vals in (motor_parameter or {}).items():
            if name not in self._MP_SLOT:
                raise KeyError(f"unknown motor parameter {name!r}")
            mp[:, self._MP_SLOT[name]] = np.broadcast_to(np.asarray(vals, dtype=np.float64), (sim.n,))
        for name, vals in (load_parameter or {}).items():
            if name not in self._LP_SLOT:
                raise KeyError(f"unknown load parameter {name!r}")
            lp[:, self._LP_SLOT[name]] = np.broadcast_to(np.asarray(vals, dtype=np.float64), (sim.n,))
        sim.set_env_params(mp, lp)

    def set_reference(self, values):
        """Push reference values [N, n_ref] for ExternalReferenceGenerator slots (used by the next step's reward)."""
        self._ensure_sim().set_reference(values)

    def state_dict(self):
        return self._ensure_sim().state_dict()

    def load_state_dict(self, sd):
        self._ensure_sim().load_state_dict(sd)

    def render(self, *_, **__):
        for v in self._visualizations:
            v.render()

    def close(self):
        self._call_callbacks("on_close")
        self._reward_function.close()
        self._reference_generator.close()
        if self._sim is not None:
            self._sim.close()
            self._sim = None
        self._physical_system.attach(None)
