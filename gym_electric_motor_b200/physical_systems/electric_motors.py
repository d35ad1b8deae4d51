"""Electric-motor descriptors: the reference's motor classes reduced to what a device build needs — the parameter
dictionaries, the limit / nominal-value derivation and the constant initial state.  The ODE right-hand sides themselves
run in CUDA (csrc/gemb200_kernels.cuh, struct Model<>); the physical parameters are handed to the library, which
derives the model constants (csrc/gemb200.cu: derive_model).

Class names, constructor kwargs, default parameters and the `limits` / `nominal_values` / `motor_parameter` properties
mirror reference physical_systems/electric_motors/*.py; the golden table tests/golden/env_table.json (dumped from the
running reference) pins every derived number.
"""
import math

import numpy as np

from .. import _cabi as K
from ..utils import update_parameter_dict

_DEFAULT_INITIALIZER = {"states": {}, "interval": None, "random_init": None, "random_params": (None, None)}


class ElectricMotor:
    """Base descriptor (reference electric_motor.py:9-326)."""

    KIND = None
    HAS_JACOBIAN = True
    CURRENTS = []
    VOLTAGES = []
    CURRENTS_IDX = []
    #: names of the motor's ODE states in solver order (after the mechanical states)
    ODE_STATES = []
    _default_motor_parameter = {}
    _default_nominal_values = {}
    _default_limits = {}
    _default_initializer = _DEFAULT_INITIALIZER

    def __init__(self, motor_parameter=None, nominal_values=None, limit_values=None, motor_initializer=None, initial_limits=None):
        self._motor_parameter = update_parameter_dict(self._default_motor_parameter, motor_parameter or {})
        self._limits = update_parameter_dict(self._default_limits, limit_values or {})
        self._nominal_values = update_parameter_dict(self._default_nominal_values, nominal_values or {})
        self._initializer = update_parameter_dict(self._default_initializer, motor_initializer or {})
        if self._initializer.get("random_init") not in (None, "uniform", "normal", "gaussian"):
            raise NotImplementedError(f"random_init={self._initializer.get('random_init')!r} (electric_motor.py:233-262 knows uniform / normal / gaussian)")
        # Key order matters (see initial_ode_state): the reference starts from the USER's `states` dict — which replaces the default one as a
        # whole, electric_motor.py:133-137 — and `initialize` then adds the missing default keys behind it (:204-206).
        defaults = self._default_initializer["states"]
        given = self._initializer["states"] or {}
        unknown = set(given) - set(defaults)
        if unknown:
            raise KeyError(f"unknown initial states {unknown}")
        self._initial_states = {k: given[k] for k in given}
        for k, v in defaults.items():
            self._initial_states.setdefault(k, v)

    # ----- reference-compatible read-only views
    @property
    def motor_parameter(self):
        return self._motor_parameter

    @property
    def limits(self):
        return self._limits

    @property
    def nominal_values(self):
        return self._nominal_values

    @property
    def initializer(self):
        return self._initializer

    def initial_ode_state(self):
        """Constant initial motor ODE state (reference reset(): `np.asarray(list(self._initial_states.values()))`,
        electric_motor.py:283-284 / synchronous_motor.py:125-131).

        Reference quirk kept on purpose: the values are taken in the KEY ORDER of the initial-state dict, not by name — the user's keys
        in the user's order, then the default keys that were not given.  The synchronous motors' default dict is ordered
        (i_sq, i_sd, epsilon) while the ODE state is (i_sd, i_sq, epsilon): `states=dict(i_sq=20)` alone therefore initialises the
        d-current, `states=dict(i_sd=-10, i_sq=20, epsilon=1)` lands where the names say (tests/golden/pmsm_cc_custom_rk4.npz, and the
        reference-vs-oracle trajectories of tests/agent_surface/kwargs_matrix_harness.py)."""
        return np.array([float(v) for v in self._initial_states.values()])

    @property
    def random_init(self):
        return self._initializer.get("random_init") in ("uniform", "normal", "gaussian")

    @property
    def gaussian_init(self):
        return self._initializer.get("random_init") in ("normal", "gaussian")

    def gaussian_params(self, lower, upper):
        """(mue, sigma) per initial state of the truncated normal (electric_motor.py:245-248): a given scalar, else the middle of
        the interval / 1."""
        rp = self._initializer.get("random_params") or (None, None)
        lower, upper = np.asarray(lower, dtype=float), np.asarray(upper, dtype=float)
        mue = np.broadcast_to(rp[0] or (upper - lower) / 2 + lower, lower.shape).astype(float)
        sigma = np.broadcast_to(rp[1] or 1, lower.shape).astype(float)
        return mue, sigma

    def initial_bounds(self, state_low, state_positions):
        """(lower, upper) per initial state in the initializer's key order (electric_motor.py:214-232): upper = the motor's
        nominal value of the state, lower = upper * state_space.low, both clipped to `interval` when given."""
        keys = list(self._initial_states)
        upper = np.array([float(self._nominal_values[k]) for k in keys])
        lower = upper * np.array([float(state_low[state_positions[k]]) for k in keys])
        interval = self._initializer.get("interval")
        if interval is not None:
            iv = np.asarray(interval, dtype=float)
            lower = np.clip(lower, a_min=iv.T[0], a_max=None)
            upper = np.clip(upper, a_min=None, a_max=iv.T[1])
        return lower, upper

    def check_initial_state(self, nominal_state, state_low, state_positions):
        """ElectricMotor.initialize constant branch (electric_motor.py:255-266): the value has to lie inside
        [nominal*low, nominal]."""
        if self.random_init:
            return
        for name, val in self._initial_states.items():
            if name not in state_positions:
                continue
            idx = state_positions[name]
            upper = nominal_state[idx]
            lower = upper * state_low[idx]
            if not (lower <= val <= upper):
                raise Exception("Initialization value has to be within nominal boundaries")

    # ----- limit derivation (electric_motor.py:297-316)
    def _base_update_limits(self, limits_d=None, nominal_d=None):
        limits_d = dict(limits_d or {})
        nominal_d = dict(nominal_d or {})
        limits_d["omega"] = self._default_limits["omega"]
        for qty, lim in limits_d.items():
            if self._limits.get(qty, 0) == 0:
                self._limits[qty] = lim
        for entry in list(self._limits.keys()):
            if self._nominal_values.get(entry, 0) == 0:
                self._nominal_values[entry] = nominal_d.get(entry, self._limits[entry])

    def torque(self, currents):
        raise NotImplementedError

    def fill_config(self, cfg):
        cfg.motor_kind = self.KIND
        slots = dict(p=K.MP_P, r_s=K.MP_R_S, l_d=K.MP_L_D, l_q=K.MP_L_Q, psi_p=K.MP_PSI_P, j_rotor=K.MP_J_ROTOR, r_a=K.MP_R_A,
                     l_a=K.MP_L_A, psi_e=K.MP_PSI_E, r_e=K.MP_R_E, l_e=K.MP_L_E, l_e_prime=K.MP_L_E_PRIME, l_m=K.MP_L_M,
                     k=K.MP_K, l_sigs=K.MP_L_SIGS, l_sigr=K.MP_L_SIGR, r_r=K.MP_R_E)
        for name, val in self._motor_parameter.items():
            if name in slots:
                cfg.motor_param[slots[name]] = float(val)


# ---------------------------------------------------------------------------------------------------------------- DC
class DcMotor(ElectricMotor):
    """reference dc_motor.py (base of the DC family, equals the externally excited motor)."""

    KIND = K.MOTOR_EXTEX_DC
    CURRENTS = ["i_a", "i_e"]
    VOLTAGES = ["u_a", "u_e"]
    CURRENTS_IDX = [0, 1]
    I_A_IDX, I_E_IDX = 0, 1  # positions in the motor's own ODE state (dc_motor.py)
    ODE_STATES = ["i_a", "i_e"]
    _default_motor_parameter = {"r_a": 16e-3, "r_e": 16e-2, "l_a": 19e-6, "l_e_prime": 1.7e-3, "l_e": 5.4e-3, "j_rotor": 0.0025}
    _default_nominal_values = dict(omega=300, torque=16.0, i=97, i_a=97, i_e=97, u=60, u_a=60, u_e=60)
    _default_limits = dict(omega=400, torque=38.0, i=210, i_a=210, i_e=210, u=60, u_a=60, u_e=60)
    _default_initializer = {"states": {"i_a": 0.0, "i_e": 0.0}, "interval": None, "random_init": None, "random_params": (None, None)}

    def __init__(self, motor_parameter=None, nominal_values=None, limit_values=None, motor_initializer=None):
        super().__init__(motor_parameter, nominal_values, limit_values, motor_initializer)
        self._update_limits()

    def torque(self, currents):
        return self._motor_parameter["l_e_prime"] * currents[0] * currents[1]

    def _dc_update_limits(self, limits_d=None):  # dc_motor.py:152-160
        limits_d = dict(limits_d or {})
        limits_d["torque"] = self.torque([self._limits[state] for state in self.CURRENTS])
        self._base_update_limits(limits_d)

    def _update_limits(self):
        self._dc_update_limits()

    def get_state_space(self, input_currents, input_voltages):  # dc_motor.py:130-150
        a, e = 0, 1
        low = {
            "omega": -1 if input_voltages.low[a] == -1 or input_voltages.low[e] == -1 else 0,
            "torque": -1 if input_currents.low[a] == -1 or input_currents.low[e] == -1 else 0,
            "i_a": -1 if input_currents.low[a] == -1 else 0,
            "i_e": -1 if input_currents.low[e] == -1 else 0,
            "u_a": -1 if input_voltages.low[a] == -1 else 0,
            "u_e": -1 if input_voltages.low[e] == -1 else 0,
        }
        high = {"omega": 1, "torque": 1, "i_a": 1, "i_e": 1, "u_a": 1, "u_e": 1}
        return low, high


class DcExternallyExcitedMotor(DcMotor):
    """reference dc_externally_excited_motor.py"""

    def _update_limits(self):
        r_a = 1 if self._motor_parameter["r_a"] == 0 else self._motor_parameter["r_a"]
        self._dc_update_limits({
            "u_a": self._default_limits["u"],
            "u_e": self._default_limits["u"],
            "i_a": self._limits.get("i", None) or self._limits["u"] / r_a,
            "i_e": self._limits.get("i", None) or self._limits["u"] / self._motor_parameter["r_e"],
        })


class DcPermanentlyExcitedMotor(DcMotor):
    """reference dc_permanently_excited_motor.py"""

    KIND = K.MOTOR_PERMEX_DC
    CURRENTS = ["i"]
    VOLTAGES = ["u"]
    CURRENTS_IDX = [0]
    I_IDX = 0
    ODE_STATES = ["i"]
    _default_motor_parameter = {"r_a": 16e-3, "l_a": 19e-6, "psi_e": 0.165, "j_rotor": 0.025}
    _default_nominal_values = dict(omega=300, torque=16.0, i=97, u=60)
    _default_limits = dict(omega=400, torque=38.0, i=210, u=60)
    _default_initializer = {"states": {"i": 0.0}, "interval": None, "random_init": None, "random_params": (None, None)}

    def torque(self, state):
        return self._motor_parameter["psi_e"] * state[0]

    def _update_limits(self):
        r_a = 1 if self._motor_parameter["r_a"] == 0 else self._motor_parameter["r_a"]
        self._dc_update_limits({"u": self._default_limits["u"], "i": self._limits["u"] / r_a})

    def get_state_space(self, input_currents, input_voltages):
        low = {
            "omega": -1 if input_voltages.low[0] == -1 else 0,
            "torque": -1 if input_currents.low[0] == -1 else 0,
            "i": -1 if input_currents.low[0] == -1 else 0,
            "u": -1 if input_voltages.low[0] == -1 else 0,
        }
        return low, {"omega": 1, "torque": 1, "i": 1, "u": 1}


class DcSeriesMotor(DcMotor):
    """reference dc_series_motor.py"""

    KIND = K.MOTOR_SERIES_DC
    CURRENTS = ["i"]
    VOLTAGES = ["u"]
    CURRENTS_IDX = [0]
    I_IDX = 0
    ODE_STATES = ["i"]
    _default_motor_parameter = {"r_a": 16e-3, "r_e": 48e-3, "l_a": 19e-6, "l_e_prime": 1.7e-3, "l_e": 5.4e-3, "j_rotor": 0.0025}
    _default_initializer = {"states": {"i": 0.0}, "interval": None, "random_init": None, "random_params": (None, None)}

    def torque(self, currents):
        return self._motor_parameter["l_e_prime"] * currents[0] * currents[0]

    def _update_limits(self):
        r_a = 1 if self._motor_parameter["r_a"] == 0 else self._motor_parameter["r_a"]
        self._dc_update_limits({"u": self._default_limits["u"], "i": self._limits["u"] / (r_a + self._motor_parameter["r_e"])})

    def get_state_space(self, input_currents, input_voltages):
        low = {"omega": 0, "torque": 0, "i": -1 if input_currents.low[0] == -1 else 0, "u": -1 if input_voltages.low[0] == -1 else 0}
        return low, {"omega": 1, "torque": 1, "i": 1, "u": 1}


class DcShuntMotor(DcMotor):
    """reference dc_shunt_motor.py"""

    KIND = K.MOTOR_SHUNT_DC
    VOLTAGES = ["u"]
    _default_motor_parameter = {"r_a": 16e-3, "r_e": 4e-1, "l_a": 19e-6, "l_e_prime": 1.7e-3, "l_e": 5.4e-3, "j_rotor": 0.0025}

    def _update_limits(self):
        r_a = 1 if self._motor_parameter["r_a"] == 0 else self._motor_parameter["r_a"]
        self._dc_update_limits({
            "u": self._default_limits["u"],
            "i_a": self._limits.get("i", None) or self._limits["u"] / r_a,
            "i_e": self._limits.get("i", None) or self._limits["u"] / self._motor_parameter["r_e"],
        })

    def get_state_space(self, input_currents, input_voltages):
        low = {
            "omega": 0,
            "torque": -1 if input_currents.low[0] == -1 else 0,
            "i_a": -1 if input_currents.low[0] == -1 else 0,
            "i_e": -1 if input_currents.low[0] == -1 else 0,
            "u": -1 if input_voltages.low[0] == -1 else 0,
        }
        return low, {"omega": 1, "torque": 1, "i_a": 1, "i_e": 1, "u": 1}


# ------------------------------------------------------------------------------------------------------ three phase
class ThreePhaseMotor(ElectricMotor):
    """reference three_phase_motor.py (limit handling :127-133).  The Clarke/Park transforms of the STEP run inside the kernel; the
    host-side versions below exist for agents (field-oriented controllers call `motor.t_32(motor.q(u_dq, eps))`, three_phase_motor.py:18-88)."""

    IO_VOLTAGES = []
    IO_CURRENTS = []

    _SQRT3_2 = 0.5 * np.sqrt(3.0)

    @staticmethod
    def t_23(quantities):
        """Clarke: (a, b, c) -> (alpha, beta), amplitude invariant"""
        a, b, c = quantities
        return np.array([(2.0 * a - b - c) / 3.0, (b - c) / np.sqrt(3.0)])

    @staticmethod
    def t_32(quantities):
        """inverse Clarke: (alpha, beta) -> (a, b, c)"""
        al, be = quantities
        h = ThreePhaseMotor._SQRT3_2 * be
        return np.array([al, -0.5 * al + h, -0.5 * al - h])

    @staticmethod
    def q(quantities, epsilon):
        """Park rotation dq -> alpha-beta by the electrical angle"""
        c, s = np.cos(epsilon), np.sin(epsilon)
        return c * quantities[0] - s * quantities[1], s * quantities[0] + c * quantities[1]

    @staticmethod
    def q_inv(quantities, epsilon):
        """alpha-beta -> dq"""
        return ThreePhaseMotor.q(quantities, -epsilon)

    def q_me(self, quantities, epsilon):
        """dq -> alpha-beta with the MECHANICAL angle (three_phase_motor.py:77-88)"""
        return self.q(quantities, epsilon * self._motor_parameter["p"])

    def _torque_limit(self):
        raise NotImplementedError

    def _three_phase_update_limits(self):
        """SynchronousMotor._update_limits (synchronous_motor.py:174-189) / SCIM (squirrel_cage_induction_motor.py:131-144)
        followed by ThreePhaseMotor._update_limits (three_phase_motor.py:127-133)."""
        voltage_limit = 0.5 * self._limits["u"]
        voltage_nominal = 0.5 * self._nominal_values["u"]
        limits_agenda, nominal_agenda = {}, {}
        for u, i in zip(self.IO_VOLTAGES, self.IO_CURRENTS):
            limits_agenda[u] = voltage_limit
            nominal_agenda[u] = voltage_nominal
            limits_agenda[i] = self._limits.get("i", None) or self._limits[u] / self._motor_parameter["r_s"]
            nominal_agenda[i] = self._nominal_values.get("i", None) or self._nominal_values[u] / self._motor_parameter["r_s"]
        self._base_update_limits(limits_agenda, nominal_agenda)
        self._base_update_limits(dict(torque=self._torque_limit()))


class SynchronousMotor(ThreePhaseMotor):
    """reference synchronous_motor.py"""

    CURRENTS = ["i_sd", "i_sq"]
    VOLTAGES = ["u_sd", "u_sq"]
    CURRENTS_IDX = [0, 1]
    I_SD_IDX, I_SQ_IDX, EPSILON_IDX = 0, 1, 2  # positions in the motor's own ODE state (synchronous_motor.py)
    ODE_STATES = ["i_sd", "i_sq", "epsilon"]
    IO_VOLTAGES = ["u_a", "u_b", "u_c", "u_sd", "u_sq"]
    IO_CURRENTS = ["i_a", "i_b", "i_c", "i_sd", "i_sq"]
    _default_initializer = {"states": {"i_sq": 0.0, "i_sd": 0.0, "epsilon": 0.0}, "interval": None, "random_init": None,
                            "random_params": (None, None)}

    def __init__(self, motor_parameter=None, nominal_values=None, limit_values=None, motor_initializer=None):
        super().__init__(motor_parameter, nominal_values or {}, limit_values or {}, motor_initializer)
        self._three_phase_update_limits()


class PermanentMagnetSynchronousMotor(SynchronousMotor):
    """reference permanent_magnet_synchronous_motor.py"""

    KIND = K.MOTOR_PMSM
    _default_motor_parameter = {"p": 3, "l_d": 0.37e-3, "l_q": 1.2e-3, "j_rotor": 0.03883, "r_s": 18e-3, "psi_p": 66e-3}
    _default_limits = dict(omega=4e3 * np.pi / 30, torque=0.0, i=400, epsilon=math.pi, u=300)
    _default_nominal_values = dict(omega=3e3 * np.pi / 30, torque=0.0, i=240, epsilon=math.pi, u=300)

    def torque(self, currents):
        mp = self._motor_parameter
        return 1.5 * mp["p"] * (mp["psi_p"] + (mp["l_d"] - mp["l_q"]) * currents[0]) * currents[1]

    def _torque_limit(self):  # :121-132
        mp = self._motor_parameter
        if mp["l_d"] == mp["l_q"]:
            return self.torque([0, self._limits["i_sq"], 0])
        i_n = self.nominal_values["i"]
        _p = mp["psi_p"] / (2 * (mp["l_d"] - mp["l_q"]))
        _q = -(i_n**2) / 2
        i_d_opt = -_p / 2 - np.sqrt((_p / 2) ** 2 - _q)
        i_q_opt = np.sqrt(i_n**2 - i_d_opt**2)
        return self.torque([i_d_opt, i_q_opt, 0])


class SynchronousReluctanceMotor(SynchronousMotor):
    """reference synchronous_reluctance_motor.py"""

    KIND = K.MOTOR_SYNRM
    _default_motor_parameter = {"p": 4, "l_d": 10.1e-3, "l_q": 4.1e-3, "j_rotor": 0.8e-3, "r_s": 0.57}
    _default_nominal_values = {"i": 10, "torque": 0, "omega": 3e3 * np.pi / 30, "epsilon": np.pi, "u": 80}
    _default_limits = {"i": 18, "torque": 0, "omega": 4.3e3 * np.pi / 30, "epsilon": np.pi, "u": 80}

    def torque(self, currents):
        mp = self._motor_parameter
        return 1.5 * mp["p"] * ((mp["l_d"] - mp["l_q"]) * currents[0]) * currents[1]

    def _torque_limit(self):  # :133-135
        return self.torque([self._limits["i_sd"] / np.sqrt(2), self._limits["i_sq"] / np.sqrt(2), 0])


class ExternallyExcitedSynchronousMotor(SynchronousMotor):
    """reference externally_excited_synchronous_motor.py"""

    KIND = K.MOTOR_EESM
    CURRENTS = ["i_sd", "i_sq", "i_e"]
    VOLTAGES = ["u_sd", "u_sq", "u_e"]
    CURRENTS_IDX = [0, 1, 2]
    I_SD_IDX, I_SQ_IDX, I_E_IDX, EPSILON_IDX = 0, 1, 2, 3
    ODE_STATES = ["i_sd", "i_sq", "i_e", "epsilon"]
    IO_VOLTAGES = ["u_a", "u_b", "u_c", "u_sd", "u_sq", "u_e"]
    IO_CURRENTS = ["i_a", "i_b", "i_c", "i_sd", "i_sq", "i_e"]
    _default_motor_parameter = {"p": 3, "l_d": 1.66e-3, "l_q": 0.35e-3, "l_m": 1.589e-3, "l_e": 1.74e-3, "j_rotor": 0.3883,
                                "r_s": 15.55e-3, "r_e": 7.2e-3, "k": 65.21}
    _default_limits = dict(omega=12e3 * np.pi / 30, torque=0.0, i=150, i_e=150, epsilon=math.pi, u=320)
    _default_nominal_values = dict(omega=4.3e3 * np.pi / 30, torque=0.0, i=120, i_e=150, epsilon=math.pi, u=320)
    _default_initializer = {"states": {"i_sq": 0.0, "i_sd": 0.0, "i_e": 0.0, "epsilon": 0.0}, "interval": None,
                            "random_init": None, "random_params": (None, None)}

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        # the reference publishes the rotor quantities referred to the stator side in the same dict (agents read e.g. mp["sigma"]),
        # externally_excited_synchronous_motor.py:125-136; the kernel derives its coefficients from the primary keys in gemb200.cu
        mp = self._motor_parameter
        turns = mp["k"]
        mp["r_E"], mp["l_M"], mp["l_E"] = 1.5 * turns**2 * mp["r_e"], 1.5 * turns * mp["l_m"], 1.5 * turns**2 * mp["l_e"]
        mp["i_k_rs"] = 2 / 3 / turns
        mp["sigma"] = 1 - mp["l_M"] ** 2 / (mp["l_d"] * mp["l_E"])

    def _derived(self):
        mp = self._motor_parameter
        return mp["k"] * 3 / 2 * mp["l_m"], 2 / 3 / mp["k"]  # l_M, i_k_rs (:129-135)

    def torque(self, currents):
        mp = self._motor_parameter
        l_M, i_k_rs = self._derived()
        return 1.5 * mp["p"] * (l_M * currents[2] * i_k_rs + (mp["l_d"] - mp["l_q"]) * currents[0]) * currents[1]

    def _torque_limit(self):  # :187-198
        mp = self._motor_parameter
        l_M, _ = self._derived()
        if mp["l_d"] == mp["l_q"]:
            return self.torque([0, self._limits["i_sq"], self._limits["i_e"], 0])
        i_n = self.nominal_values["i"]
        _p = l_M * i_n / (2 * (mp["l_d"] - mp["l_q"]))
        _q = -(i_n**2) / 2
        if mp["l_d"] < mp["l_q"]:
            i_d_opt = -_p / 2 - np.sqrt((_p / 2) ** 2 - _q)
        else:
            i_d_opt = -_p / 2 + np.sqrt((_p / 2) ** 2 - _q)
        i_q_opt = np.sqrt(i_n**2 - i_d_opt**2)
        return self.torque([i_d_opt, i_q_opt, self._limits["i_e"], 0])


class InductionMotor(ThreePhaseMotor):
    """reference induction_motor.py"""

    CURRENTS = ["i_salpha", "i_sbeta"]
    FLUXES = ["psi_ralpha", "psi_rbeta"]
    VOLTAGES = ["u_salpha", "u_sbeta"]
    CURRENTS_IDX = [0, 1]
    I_SALPHA_IDX, I_SBETA_IDX, PSI_RALPHA_IDX, PSI_RBETA_IDX, EPSILON_IDX = 0, 1, 2, 3, 4  # induction_motor.py
    FLUX_IDX = [2, 3]
    STATOR_VOLTAGES = ["u_salpha", "u_sbeta"]
    ODE_STATES = ["i_salpha", "i_sbeta", "psi_ralpha", "psi_rbeta", "epsilon"]
    IO_VOLTAGES = ["u_sa", "u_sb", "u_sc", "u_salpha", "u_sbeta", "u_sd", "u_sq"]
    IO_CURRENTS = ["i_sa", "i_sb", "i_sc", "i_salpha", "i_sbeta", "i_sd", "i_sq"]
    _default_motor_parameter = {"p": 2, "l_m": 143.75e-3, "l_sigs": 5.87e-3, "l_sigr": 5.87e-3, "j_rotor": 1.1e-3, "r_s": 2.9338,
                                "r_r": 1.355}
    _default_limits = dict(omega=4e3 * np.pi / 30, torque=0.0, i=5.5, epsilon=math.pi, u=560)
    _default_nominal_values = dict(omega=3e3 * np.pi / 30, torque=0.0, i=3.9, epsilon=math.pi, u=560)
    _default_initializer = {"states": {"i_salpha": 0.0, "i_sbeta": 0.0, "psi_ralpha": 0.0, "psi_rbeta": 0.0, "epsilon": 0.0},
                            "interval": None, "random_init": None, "random_params": (None, None)}

    def __init__(self, motor_parameter=None, nominal_values=None, limit_values=None, motor_initializer=None, initial_limits=None):
        super().__init__(motor_parameter, nominal_values, limit_values, motor_initializer, initial_limits)
        self._three_phase_update_limits()

    def torque(self, states):
        mp = self._motor_parameter
        return 1.5 * mp["p"] * mp["l_m"] / (mp["l_m"] + mp["l_sigr"]) * (states[2] * states[1] - states[3] * states[0])

    def _torque_limit(self):  # :219-234
        mp = self._motor_parameter
        return 1.5 * mp["p"] * mp["l_m"] ** 2 / (mp["l_m"] + mp["l_sigr"]) * self._limits["i_sd"] * self._limits["i_sq"] / 2

    # ----- initial states (induction_motor.py:178-185, :250-285; squirrel_cage_induction_motor.py:146-157; doubly_fed_induction_motor.py:154-165)
    def flux_limit_constants(self):
        """The constants of InductionMotor._flux_limit, for the device (gemb200.h: init_im): psi_d_max at omega = 0, the three numerator
        terms and the denominator factor of the omega != 0 branch, l_m."""
        mp = self._motor_parameter
        l_s, l_r = mp["l_m"] + mp["l_sigs"], mp["l_m"] + mp["l_sigr"]
        l_mr = mp["l_m"] / l_r
        sigma = (l_s * l_r - mp["l_m"] ** 2) / (l_s * l_r)
        u_q = float(self._nominal_values["u_sq"]) + l_mr * float(self._nominal_values.get("u_rq", 0.0) if self.KIND == K.MOTOR_DFIM else 0.0)
        return [mp["l_m"] * float(self._nominal_values["i_sd"]), mp["p"] * sigma * l_s, mp["r_s"] + mp["r_r"] * l_mr**2, u_q, mp["p"] * l_mr, mp["l_m"], 0.0, 0.0]

    def initial_bounds(self, state_low, state_positions):
        """ElectricMotor.initialize, induction branch (electric_motor.py:197-213): upper = |initial limit| of the state, lower = -upper.
        The two flux limits are re-derived per reset on the device (flux_limit_constants); here they only carry the user's `interval`."""
        keys = list(self._initial_states)
        big = 1e30
        upper = np.array([big if k in self.FLUXES else abs(float(self._nominal_values[k])) for k in keys])
        lower = -upper
        interval = self._initializer.get("interval")
        if interval is not None:
            iv = np.asarray(interval, dtype=float)
            lower = np.clip(lower, a_min=iv.T[0], a_max=None)
            upper = np.clip(upper, a_min=None, a_max=iv.T[1])
        return lower, upper

    def gaussian_params(self, lower, upper):
        """mue = random_params[0] or the middle of the interval — for the flux states the interval is per env and per reset, so a missing
        mue is passed on as NaN and resolved on the device (gemb200.h: init_mu)"""
        rp = self._initializer.get("random_params") or (None, None)
        lower = np.asarray(lower, dtype=float)
        mue = np.full(lower.shape, float(rp[0]) if rp[0] else np.nan)
        sigma = np.full(lower.shape, float(rp[1] or 1))
        return mue, sigma

    def check_initial_state(self, nominal_state, state_low, state_positions):
        """Constant initial states are checked against +-|initial limit| (electric_motor.py:255-266).  The reference's flux limits depend on
        a draw from the global numpy RNG, so a constant non-zero flux passes or raises at random there; here it is accepted up to the
        omega = 0 limit l_m * i_sd_nominal and refused beyond it."""
        if self.random_init:
            return
        psi_max = self.flux_limit_constants()[0]
        for name, val in self._initial_states.items():
            lim = psi_max if name in self.FLUXES else abs(float(self._nominal_values[name]))
            if not (-lim <= val <= lim):
                raise Exception("Initialization value has to be within nominal boundaries")

    def fill_init_config(self, cfg):
        if self.random_init:
            cfg.init_im_valid = 1
            for j, v in enumerate(self.flux_limit_constants()):
                cfg.init_im[j] = float(v)


class SquirrelCageInductionMotor(InductionMotor):
    """reference squirrel_cage_induction_motor.py"""

    KIND = K.MOTOR_SCIM


class DoublyFedInductionMotor(InductionMotor):
    """reference doubly_fed_induction_motor.py: the induction model with accessible rotor windings (second B6 bridge)."""

    KIND = K.MOTOR_DFIM
    ROTOR_VOLTAGES = ["u_ralpha", "u_rbeta"]
    ROTOR_CURRENTS = ["i_ralpha", "i_rbeta"]
    IO_VOLTAGES = InductionMotor.IO_VOLTAGES + ["u_ra", "u_rb", "u_rc", "u_rd", "u_rq"]
    IO_CURRENTS = InductionMotor.IO_CURRENTS + ["i_ra", "i_rb", "i_rc", "i_rd", "i_rq"]
    _default_motor_parameter = {"p": 2, "l_m": 297.5e-3, "l_sigs": 25.71e-3, "l_sigr": 25.71e-3, "j_rotor": 13.695e-3, "r_s": 4.42, "r_r": 3.51}
    _default_limits = dict(omega=1800 * np.pi / 30, torque=0.0, i=9, epsilon=math.pi, u=720)
    _default_nominal_values = dict(omega=1650 * np.pi / 30, torque=0.0, i=7.5, epsilon=math.pi, u=720)

    def _three_phase_update_limits(self):
        """DoublyFedInductionMotor._update_limits (doubly_fed_induction_motor.py:126-147): rotor quantities included, current fallback
        u / r_r; then ThreePhaseMotor._update_limits."""
        voltage_limit = 0.5 * self._limits["u"]
        voltage_nominal = 0.5 * self._nominal_values["u"]
        limits_agenda, nominal_agenda = {}, {}
        for u, i in zip(self.IO_VOLTAGES + self.ROTOR_VOLTAGES, self.IO_CURRENTS + self.ROTOR_CURRENTS):
            limits_agenda[u] = voltage_limit
            nominal_agenda[u] = voltage_nominal
            limits_agenda[i] = self._limits.get("i", None) or self._limits[u] / self._motor_parameter["r_r"]
            nominal_agenda[i] = self._nominal_values.get("i", None) or self._nominal_values[u] / self._motor_parameter["r_r"]
        self._base_update_limits(limits_agenda, nominal_agenda)
        self._base_update_limits(dict(torque=self._torque_limit()))

