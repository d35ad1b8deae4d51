"""Batched SCML (Supply-Converter-Motor-Load) physical systems — the host-side mirror of reference
physical_systems/physical_systems.py.  A system object derives everything the reference derives on the host
(state names, positions, limits, nominal state, state/action spaces) and owns the device handle that steps all N
environments; `simulate` / `reset` keep the reference's names and return the NORMALISED state for all envs.

Layout on the device and the kernels are described in DESIGN.md; the constructor keeps the reference's signature
(`converter, motor, load, supply, ode_solver, tau, calc_jacobian`) plus the batch arguments `num_envs, device, dtype`.
"""
import numpy as np

from .. import _cabi as K
from ..spaces import Box
from ..utils import set_state_array
from .converters import ContDynamicallyAveragedConverter, FiniteConverter, PowerElectronicConverter
from .electric_motors import (DcExternallyExcitedMotor, DcMotor, DcPermanentlyExcitedMotor, DcSeriesMotor, DcShuntMotor, DoublyFedInductionMotor, ElectricMotor,
                              ExternallyExcitedSynchronousMotor, InductionMotor, SynchronousMotor)
from .mechanical_loads import MechanicalLoad
from .solvers import OdeSolver
from .voltage_supplies import IdealVoltageSupply, VoltageSupply


class PhysicalSystem:
    """reference core.py:589-705 (batched: state arrays have a leading env dimension)."""

    def __init__(self, action_space, state_space, state_names, tau):
        self._action_space = action_space
        self._state_space = state_space
        self._state_names = list(state_names)
        self._state_positions = {key: index for index, key in enumerate(self._state_names)}
        self._tau = tau
        self._k = 0

    @property
    def tau(self):
        return self._tau

    @property
    def k(self):
        return self._k

    @property
    def state_names(self):
        return self._state_names

    @property
    def state_positions(self):
        return self._state_positions

    @property
    def action_space(self):
        return self._action_space

    @property
    def state_space(self):
        return self._state_space

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class SCMLSystem(PhysicalSystem):
    """reference physical_systems.py:13-287, batched.

    Extra args: num_envs (N), device (CUDA ordinal or 'cuda:k'), dtype ('float32' = fp32 state + double-float rotor angle,
    'float64'), layout ('aos' row-per-env [N, n_state] | 'soa' field-major [n_state, N])."""

    # class-level defaults of the index attributes, as in the reference (physical_systems.py:20-24); instances overwrite them in _set_indices
    OMEGA_IDX = 0
    TORQUE_IDX = 1
    CURRENTS_IDX = []
    VOLTAGES_IDX = []
    U_SUP_IDX = -1

    _MOTOR_BASE = ElectricMotor

    def __init__(self, converter, motor, load, supply, ode_solver, tau=1e-4, calc_jacobian=None, num_envs=1, device=0,
                 dtype="float32", layout="aos", env_index_offset=0, **_):
        for obj, base, what in ((converter, PowerElectronicConverter, "converter"), (motor, self._MOTOR_BASE, "motor"),
                                (load, MechanicalLoad, "load"), (supply, VoltageSupply, "supply"), (ode_solver, OdeSolver, "ode_solver")):
            if not isinstance(obj, base):
                raise TypeError(f"{what}={type(obj).__name__} is not a built-in {base.__name__} of gym_electric_motor_b200; "
                                "user-defined Python components cannot run inside the CUDA kernel (INTEGRATION.md)")
        self._converter, self._electrical_motor, self._mechanical_load, self._supply, self._ode_solver = converter, motor, load, supply, ode_solver
        self.num_envs = int(num_envs)
        self._device = _device_index(device)
        self._dtype = K.F32 if str(dtype).replace("torch.", "") in ("float32", "f32") else K.F64
        self._layout = K.LAYOUT_SOA if str(layout).lower() == "soa" else K.LAYOUT_AOS
        self._env_index_offset = int(env_index_offset)
        self._mechanical_load.set_j_rotor(self._electrical_motor.motor_parameter["j_rotor"])  # :83
        state_names = self._build_state_names()
        self._set_indices(state_names)
        state_space = self._build_state_space(state_names)
        super().__init__(self._converter.action_space, state_space, state_names, tau)
        self._limits = np.zeros(len(state_names))
        self._nominal_state = np.zeros(len(state_names))
        self._set_limits()
        self._set_nominal_state()
        self._converter.tau = self.tau  # :103
        if self._converter.interlocking_time < 0 or self._converter.interlocking_time >= self.tau:
            raise ValueError("interlocking_time must be within [0, tau)")
        self._electrical_motor.check_initial_state(self._nominal_state, self._state_space.low, self._state_positions)
        self._mechanical_load.check_initial_state(self._nominal_state, self._state_space.low, self._state_positions)
        self._sim = None
        self._owns_sim = False
        self._action_dq, self._angle_advance, self._dead_steps, self._dead_outer = 0, 0.0, 0, 0
        self._state_ops = []
        self._base_limits = self._limits.copy()  # limits of the system's own state vector (normalisation inside the kernel)

    def apply_wrappers(self, wrappers):
        """Fuse reference-style physical_system_wrappers into the kernel configuration (see physical_system_wrappers.py).
        The list order has the reference's meaning: each entry wraps the system built from the previous ones."""
        from ..physical_system_wrappers import (CosSinProcessor, CurrentSumProcessor, DeadTimeProcessor, DqToAbcActionProcessor, FluxObserver,
                                                StateNoiseProcessor)
        from .converters import FiniteConverter
        from .electric_motors import DcShuntMotor, DoublyFedInductionMotor, ExternallyExcitedSynchronousMotor, InductionMotor, SynchronousMotor

        for w in wrappers:
            if isinstance(w, CurrentSumProcessor):
                if isinstance(self._electrical_motor, DcShuntMotor) and w._currents == ("i_a", "i_e") and w._limit == "max" and "i_sum" in self._state_names \
                        and not self._state_ops:
                    continue  # the shunt system's own i_sum state (the reference's ShuntDc envs wrap the system with exactly this processor)
                self._add_state_op(w)
            elif isinstance(w, DeadTimeProcessor):
                if self._dead_steps:
                    raise NotImplementedError("only one DeadTimeProcessor is supported")
                self._dead_steps = w.dead_time
                self._dead_outer = 1 if self._action_dq else 0  # it wraps an existing dq transformation -> queue of dq actions
            elif isinstance(w, DqToAbcActionProcessor):
                from .electric_motors import ThreePhaseMotor

                assert isinstance(self._electrical_motor, ThreePhaseMotor), \
                    "The motor in the system has to derive from the ThreePhaseMotor to define transformations."  # dq_to_abc_action_processor.py:59-60
                if w.angle_name not in self._state_names:  # the reference runs state_names.index(angle_name) first (:63): its ValueError
                    raise ValueError(f"{w.angle_name!r} is not in list")
                dfim = isinstance(self._electrical_motor, DoublyFedInductionMotor)
                scim = isinstance(self._electrical_motor, InductionMotor) and not dfim
                if dfim != bool(w.dfim):
                    raise NotImplementedError("DqToAbcActionProcessor.make('DFIM') goes with the DFIM system and only with it")
                if dfim:
                    assert "psi_angle" in self._state_names, "Angle psi_angle not in the states of the physical system. Probably a flux observer is required."
                if not isinstance(self._electrical_motor, (SynchronousMotor, InductionMotor)) or isinstance(self._converter, FiniteConverter):
                    raise NotImplementedError("DqToAbcActionProcessor needs a PMSM/SynRM/EESM/SCIM system with a continuous converter")
                if self._action_dq:
                    raise NotImplementedError("the system already takes dq actions")
                assert w.angle_name in self._state_names, (
                    f"Angle {w.angle_name} not in the states of the physical system. Probably a flux observer is required.")
                if scim != (w.angle_name == "psi_angle"):
                    raise NotImplementedError("angle 'psi_angle' goes with the induction motor, 'epsilon' with the synchronous motors")
                self._action_dq = 3 if dfim else (2 if scim else 1)
                self._angle_advance = 0.5 + self._dead_steps  # dq_to_abc_action_processor.py:69-72
                self._action_space = w.action_space(isinstance(self._electrical_motor, ExternallyExcitedSynchronousMotor))
            elif isinstance(w, (CosSinProcessor, FluxObserver, StateNoiseProcessor)):
                self._add_state_op(w)
            else:
                raise NotImplementedError(f"{type(w).__name__} is not a device-side physical-system wrapper")
        return self

    def _add_state_op(self, w):
        """State-vector wrappers: the bookkeeping of their set_physical_system (names, positions, limits, nominal state, space);
        the arithmetic runs in the kernel (gemb200_state_op)."""
        from ..physical_system_wrappers import CosSinProcessor, CurrentSumProcessor, FluxObserver
        from .electric_motors import InductionMotor

        if len(self._state_ops) >= K.MAX_STATE_OPS:
            raise NotImplementedError(f"at most {K.MAX_STATE_OPS} state-vector wrappers per system")
        names, low, high = list(self._state_names), self._state_space.low, self._state_space.high
        if isinstance(w, CosSinProcessor):  # cos_sin_processor.py:39-58
            idx = self._state_positions[w.angle]
            rm = [idx] if w._remove_angle else []
            low = np.concatenate((np.delete(low, rm), [-1.0, -1.0]))
            high = np.concatenate((np.delete(high, rm), [1.0, 1.0]))
            self._limits = np.concatenate((np.delete(self._limits, rm), [1.0, 1.0]))
            self._nominal_state = np.concatenate((np.delete(self._nominal_state, rm), [1.0, 1.0]))
            names = list(np.delete(names, rm)) + [f"cos({w.angle})", f"sin({w.angle})"]
            self._state_ops.append(dict(kind=K.SOP_COS_SIN, idx=[idx, int(w._remove_angle), 0, 0], mask=0, param=[]))
        elif isinstance(w, CurrentSumProcessor):  # current_sum_processor.py:24-44
            if "i_sum" in names:
                raise NotImplementedError("the system already has an i_sum state")
            ci = [self._state_positions[c] for c in w._currents]  # KeyError for an unknown current, like the reference
            pick = max if w._limit == "max" else np.sum
            low, high = np.concatenate((low, [-1.0])), np.concatenate((high, [1.0]))
            self._limits = np.concatenate((self._limits, [pick(self._limits[ci])]))
            self._nominal_state = np.concatenate((self._nominal_state, [pick(self._nominal_state[ci])]))
            names = names + ["i_sum"]
            mask = 0
            for j in ci:
                mask |= 1 << j
            self._state_ops.append(dict(kind=K.SOP_CURRENT_SUM, idx=[0, 0, 0, 0], mask=mask, param=[]))
        elif isinstance(w, FluxObserver):  # flux_observer.py:56-79
            assert isinstance(self._electrical_motor, InductionMotor)
            mp = self._electrical_motor.motor_parameter
            l_m, l_r, r_r, p = mp["l_m"], mp["l_m"] + mp["l_sigr"], mp["r_r"], mp["p"]
            psi_limit = l_m * self._limits[names.index("i_sd")]
            ci = [self._state_positions[n] for n in w._current_names]
            oi = self._state_positions["omega"]
            param = [r_r * l_m / l_r, r_r / l_r, p, psi_limit] + [self._limits[j] for j in ci] + [self._limits[oi]]
            low = np.concatenate((low, [-psi_limit, -np.pi]))  # (sic) the reference's space is in physical units here
            high = np.concatenate((high, [psi_limit, np.pi]))
            self._limits = np.concatenate((self._limits, [psi_limit, np.pi]))
            self._nominal_state = np.concatenate((self._nominal_state, [psi_limit, np.pi]))
            names = names + ["psi_abs", "psi_angle"]
            self._state_ops.append(dict(kind=K.SOP_FLUX_OBSERVER, idx=ci + [oi], mask=0, param=param))
        else:  # StateNoiseProcessor: state_noise_processor.py:69-72
            sel = names if (isinstance(w._states, str) and w._states == "all") else list(w._states)
            mask = 0
            for n in sel:
                mask |= 1 << self._state_positions[n]
            dist = {"normal": K.NOISE_NORMAL, "uniform": K.NOISE_UNIFORM, "laplace": K.NOISE_LAPLACE}[w._random_dist]
            self._state_ops.append(dict(kind=K.SOP_NOISE, idx=[dist, 0, 0, 0], mask=mask, param=list(w.params)))
        self._state_names = names
        self._state_positions = {key: index for index, key in enumerate(names)}
        self._state_space = Box(low, high, dtype=np.float64)

    # ------------------------------------------------------------------ reference-compatible properties
    @property
    def limits(self):
        return self._limits

    @property
    def nominal_state(self):
        return self._nominal_state

    @property
    def supply(self):
        return self._supply

    @property
    def converter(self):
        return self._converter

    @property
    def electrical_motor(self):
        return self._electrical_motor

    @property
    def mechanical_load(self):
        return self._mechanical_load

    @property
    def ode_solver(self):
        return self._ode_solver

    # ------------------------------------------------------------------ host derivations
    def _set_limits(self):  # :105-112
        for ind, state in enumerate(self._state_names):
            self._limits[ind] = min(self._electrical_motor.limits.get(state, np.inf), self._mechanical_load.limits.get(state, np.inf))
        self._limits[self._state_positions["u_sup"]] = self.supply.u_nominal

    def _set_nominal_state(self):  # :114-123
        for ind, state in enumerate(self._state_names):
            self._nominal_state[ind] = min(self._electrical_motor.nominal_values.get(state, np.inf),
                                           self._mechanical_load.nominal_values.get(state, np.inf))
        self._nominal_state[self._state_positions["u_sup"]] = self.supply.u_nominal

    def _build_state_names(self):
        raise NotImplementedError

    def _build_state_space(self, state_names):
        raise NotImplementedError

    def _set_indices(self, state_names):
        """positions in the system's own state vector (before wrappers) with the attribute names agents read
        (physical_systems.py:141-162, :462-485, :594-617, :737-763): currents are the `i*` names, voltages the `u*` names except u_sup"""
        self.OMEGA_IDX = state_names.index("omega")
        self.TORQUE_IDX = state_names.index("torque")
        self.CURRENTS_IDX = [k for k, n in enumerate(state_names) if n.startswith("i") and n != "i_sum"]
        self.VOLTAGES_IDX = [k for k, n in enumerate(state_names) if n.startswith("u") and n != "u_sup"]
        self.U_SUP_IDX = [state_names.index("u_sup")]
        if "epsilon" in state_names:
            self.EPSILON_IDX = state_names.index("epsilon")

    def initial_ode_state(self):
        """[omega, motor states...] constant initial ODE state (SCMLSystem.reset :263-270)."""
        return np.concatenate(([self._mechanical_load.initial_omega()], self._electrical_motor.initial_ode_state()))

    # ------------------------------------------------------------------ config / handle
    def fill_config(self, cfg):
        """Write the physics part of a gemb200_config (everything except constraints / reward / references)."""
        cfg.n_envs = self.num_envs
        cfg.device = self._device
        cfg.dtype = self._dtype
        cfg.layout = self._layout
        cfg.env_index_offset = self._env_index_offset
        cfg.finite = int(isinstance(self._converter, FiniteConverter))
        slots = self._converter.slots()
        for i in range(2):
            cfg.converter_kind[i] = slots[i] if i < len(slots) else K.CONV_NONE
        cfg.tau = float(self.tau)
        cfg.interlocking_time = float(self._converter.interlocking_time)
        ils = self._converter.interlocking_times() if hasattr(self._converter, "interlocking_times") else []
        cfg.interlocking_time1 = float(ils[1]) if len(ils) > 1 and ils[1] != ils[0] else -1.0
        self._supply.fill_config(cfg)
        self._electrical_motor.fill_config(cfg)
        self._ode_solver.fill_config(cfg)
        self._mechanical_load.fill_config(cfg)  # after the solver: a tabulated speed profile needs the sub-step grid
        for i, v in enumerate(self._base_limits):
            cfg.limits[i] = float(v)
        for i, v in enumerate(self.initial_ode_state()):
            cfg.init_ode[i] = float(v)
        # random initial states (uniform): bounds per ODE state; constant states get lo == hi
        m, ld = self._electrical_motor, self._mechanical_load
        if m.random_init or ld.random_init:
            cfg.init_random = 1
            init = self.initial_ode_state()
            lo, hi = init.copy(), init.copy()
            if ld.random_init:
                lo[0], hi[0] = ld.initial_bounds(self._nominal_state, self._state_space.low, self._state_positions)
            if m.random_init:
                mlo, mhi = m.initial_bounds(self._state_space.low, self._state_positions)
                lo[1:], hi[1:] = mlo, mhi  # initializer key order == reference's assignment order (see initial_ode_state)
            for i in range(len(init)):
                cfg.init_lo[i], cfg.init_hi[i] = float(lo[i]), float(hi[i])
            # truncated normal (random_init='normal' / 'gaussian'): per-state mue / sigma; degenerate intervals stay constant
            if ld.random_init and ld.gaussian_init and hi[0] > lo[0]:
                cfg.init_dist[0] = 1
                cfg.init_mu[0], cfg.init_sigma[0] = ld.gaussian_params(lo[0], hi[0])
            if m.random_init and m.gaussian_init:
                mu, sg = m.gaussian_params(lo[1:], hi[1:])
                for j in range(len(mu)):
                    if hi[1 + j] > lo[1 + j]:
                        cfg.init_dist[1 + j] = 1
                        cfg.init_mu[1 + j], cfg.init_sigma[1 + j] = float(mu[j]), float(sg[j])
            if m.random_init and hasattr(m, "fill_init_config"):
                m.fill_init_config(cfg)  # induction motors: constants of the per-reset flux limits
        cfg.action_dq = int(self._action_dq)
        cfg.angle_advance = float(self._angle_advance)
        cfg.dead_time_steps = int(self._dead_steps)
        cfg.dead_time_outer = int(self._dead_outer)
        cfg.n_state_ops = len(self._state_ops)
        for k, op in enumerate(self._state_ops):
            cfg.sop_kind[k] = op["kind"]
            cfg.sop_mask[k] = op["mask"]
            for q, v in enumerate(op["idx"]):
                cfg.sop_idx[k][q] = int(v)
            for q, v in enumerate(op["param"]):
                cfg.sop_param[k][q] = float(v)
        return cfg

    def attach(self, sim, owns=False):
        """Called by the environment: the handle that also carries the fused epilogue."""
        self._sim, self._owns_sim = sim, owns

    def _ensure_sim(self):
        if self._sim is None:
            from ..vector_sim import VectorSim

            cfg = self.fill_config(K.new_config())
            self._sim, self._owns_sim = VectorSim(cfg), True
        return self._sim

    # ------------------------------------------------------------------ PhysicalSystem API (batched)
    def reset(self, initial_state=None, mask=None):
        """Reset all (or the masked) envs; returns the normalised state [N, n_state] (reference :256-287)."""
        sim = self._ensure_sim()
        obs, _ = sim.reset(mask)
        self._k = 0
        return obs

    def simulate(self, action, *_, **__):
        """One step for all envs; returns the normalised state [N, n_state] (reference :171-203 and overrides)."""
        sim = self._ensure_sim()
        obs, _, _, _ = sim.step(action)
        self._k += 1
        return obs

    def close(self):
        if self._sim is not None and self._owns_sim:
            self._sim.close()
        self._sim = None


def _device_index(device):
    if isinstance(device, int):
        return device
    s = str(device)
    if s in ("cuda", "gpu"):
        try:
            import torch

            return torch.cuda.current_device()
        except Exception:
            return 0
    if ":" in s:
        return int(s.split(":")[1])
    return int(s)


class DcMotorSystem(SCMLSystem):
    """reference physical_systems.py:290-318"""

    _MOTOR_BASE = DcMotor

    def _build_state_names(self):
        names = self._mechanical_load.state_names + ["torque"] + self._electrical_motor.CURRENTS + self._electrical_motor.VOLTAGES + ["u_sup"]
        if isinstance(self._electrical_motor, DcShuntMotor):
            # every ShuntDc env of the reference wraps the system in CurrentSumProcessor(('i_a','i_e'))
            # (envs/gym_dcm/shunt_dc_motor_env/*.py); the i_sum state is produced natively by the kernel
            names = names + ["i_sum"]
        return names

    def _build_state_space(self, state_names):
        low, high = self._electrical_motor.get_state_space(self._converter.currents, self._converter.voltages)
        low_m, high_m = self._mechanical_load.get_state_space((low["omega"], high["omega"]))
        low.update(low_m)
        high.update(high_m)
        high["u_sup"] = self._supply.supply_range[1] / self._supply.u_nominal
        low["u_sup"] = self._supply.supply_range[0] / self._supply.u_nominal if self._supply.supply_range[0] != self._supply.supply_range[1] else 0
        if "i_sum" in state_names:
            low["i_sum"], high["i_sum"] = -1.0, 1.0  # current_sum_processor.py:30-32
        return Box(set_state_array(low, state_names), set_state_array(high, state_names), dtype=np.float64)

    def _set_limits(self):
        names = [n for n in self._state_names if n != "i_sum"]
        for ind, state in enumerate(names):
            self._limits[ind] = min(self._electrical_motor.limits.get(state, np.inf), self._mechanical_load.limits.get(state, np.inf))
        self._limits[self._state_positions["u_sup"]] = self.supply.u_nominal
        if "i_sum" in self._state_positions:  # current_sum_processor.py:34-37 (limit='max')
            idx = [self._state_positions["i_a"], self._state_positions["i_e"]]
            self._limits[self._state_positions["i_sum"]] = max(self._limits[idx])

    def _set_nominal_state(self):
        names = [n for n in self._state_names if n != "i_sum"]
        for ind, state in enumerate(names):
            self._nominal_state[ind] = min(self._electrical_motor.nominal_values.get(state, np.inf),
                                           self._mechanical_load.nominal_values.get(state, np.inf))
        self._nominal_state[self._state_positions["u_sup"]] = self.supply.u_nominal
        if "i_sum" in self._state_positions:
            idx = [self._state_positions["i_a"], self._state_positions["i_e"]]
            self._nominal_state[self._state_positions["i_sum"]] = max(self._nominal_state[idx])


class ThreePhaseMotorSystem(SCMLSystem):
    """reference physical_systems.py:321-415.  The transformations of the step run inside the kernel; the helper methods keep the
    reference's names for agents (observers, field-oriented controllers)."""

    _NAMES = []

    @staticmethod
    def _angle(epsilon_el, normed_epsilon):
        return epsilon_el * np.pi if normed_epsilon else epsilon_el

    def abc_to_alphabeta_space(self, abc_quantities):
        return self._electrical_motor.t_23(abc_quantities)

    def alphabeta_to_abc_space(self, alphabeta_quantities):
        return self._electrical_motor.t_32(alphabeta_quantities)

    def abc_to_dq_space(self, abc_quantities, epsilon_el, normed_epsilon=False):
        m = self._electrical_motor
        return m.q_inv(m.t_23(abc_quantities), self._angle(epsilon_el, normed_epsilon))

    def dq_to_abc_space(self, dq_quantities, epsilon_el, normed_epsilon=False):
        m = self._electrical_motor
        return m.t_32(m.q(dq_quantities, self._angle(epsilon_el, normed_epsilon)))

    def alphabeta_to_dq_space(self, alphabeta_quantities, epsilon_el, normed_epsilon=False):
        return self._electrical_motor.q_inv(alphabeta_quantities, self._angle(epsilon_el, normed_epsilon))

    def dq_to_alphabeta_space(self, dq_quantities, epsilon_el, normed_epsilon=False):
        return self._electrical_motor.q(dq_quantities, self._angle(epsilon_el, normed_epsilon))

    def __init__(self, control_space="abc", **kwargs):
        assert control_space in ("abc", "dq")
        self.control_space = control_space
        super().__init__(**kwargs)
        if control_space == "dq":  # physical_systems.py:423-435, :491-492 (SCIM :779-780): a_abc = T32 q(a_dq, angle), no advance
            from .converters import FiniteConverter
            from .electric_motors import ExternallyExcitedSynchronousMotor

            assert not isinstance(self._converter, FiniteConverter), "dq-control space is only available for Continuous Controlled Converters"
            if isinstance(self._electrical_motor, ExternallyExcitedSynchronousMotor):
                raise NotImplementedError("the reference's EESM system ignores control_space='dq' in simulate (physical_systems.py:619-657)")
            self._action_dq, self._angle_advance = 1, 0.0
            self._action_space = Box(-1, 1, shape=(2,), dtype=np.float64)

    def _build_state_names(self):
        return self._mechanical_load.state_names + list(self._NAMES)

    def _build_state_space(self, state_names):  # :437-442
        low = -1 * np.ones(len(state_names))
        low[state_names.index("u_sup")] = 0.0
        return Box(low, np.ones(len(state_names)), dtype=np.float64)


class SynchronousMotorSystem(ThreePhaseMotorSystem):
    """reference physical_systems.py:418-561"""

    _MOTOR_BASE = SynchronousMotor
    _NAMES = ["torque", "i_a", "i_b", "i_c", "i_sd", "i_sq", "u_a", "u_b", "u_c", "u_sd", "u_sq", "epsilon", "u_sup"]


class ExternallyExcitedSynchronousMotorSystem(SynchronousMotorSystem):
    """reference physical_systems.py:564-693"""

    _MOTOR_BASE = ExternallyExcitedSynchronousMotor
    _NAMES = ["torque", "i_a", "i_b", "i_c", "i_sd", "i_sq", "i_e", "u_a", "u_b", "u_c", "u_sd", "u_sq", "u_e", "epsilon", "u_sup"]


class SquirrelCageInductionMotorSystem(ThreePhaseMotorSystem):
    """reference physical_systems.py:696-847"""

    _MOTOR_BASE = InductionMotor
    _NAMES = ["torque", "i_sa", "i_sb", "i_sc", "i_sd", "i_sq", "u_sa", "u_sb", "u_sc", "u_sd", "u_sq", "epsilon", "u_sup"]


class DoublyFedInductionMotorSystem(ThreePhaseMotorSystem):
    """reference physical_systems.py:850-1113: stator and rotor each fed by a B6 bridge (Cont/FiniteMultiConverter of two B6)."""

    _MOTOR_BASE = DoublyFedInductionMotor
    _NAMES = ["torque", "i_sa", "i_sb", "i_sc", "i_sd", "i_sq", "i_ra", "i_rb", "i_rc", "i_rd", "i_rq", "u_sa", "u_sb", "u_sc", "u_sd", "u_sq",
              "u_ra", "u_rb", "u_rc", "u_rd", "u_rq", "epsilon", "u_sup"]

    def __init__(self, control_space="abc", **kwargs):
        if control_space != "abc":
            raise NotImplementedError("dq actions for the DFIM need the 4-action DqToAbcActionProcessor (dq_to_abc_action_processor.py:108-137); "
                                      "not on the device path")
        super().__init__(control_space=control_space, **kwargs)
        if self._converter.slots() != [K.CONV_B6, K.CONV_B6]:
            raise ValueError("the DFIM system needs a multi converter of two B6 bridges (stator, rotor)")
