"""Mechanical-load descriptors (reference physical_systems/mechanical_loads/*.py); the ODE runs in the kernel (load_ode)."""
import numpy as np

from .. import _cabi as K
from ..utils import update_parameter_dict


class MechanicalLoad:
    """reference mechanical_load.py:9-236"""

    OMEGA_IDX = 0
    HAS_JACOBIAN = True
    KIND = None
    _default_initializer = {"states": {"omega": 0.0}, "interval": None, "random_init": None, "random_params": (None, None)}

    def __init__(self, state_names=None, j_load=0.0, load_initializer=None):
        self._j_total = self._j_load = j_load
        self._state_names = list(state_names or ["omega"])
        self._limits = {}
        self._nominal_values = {}
        self._initializer = dict(self._default_initializer)
        self._initializer["states"] = dict(self._default_initializer["states"])
        li = dict(load_initializer or {})
        if "states" in li:
            li["states"] = dict(li["states"])
        self._initializer.update(li)
        if self._initializer.get("random_init") not in (None, "uniform", "normal", "gaussian"):
            raise NotImplementedError(f"random_init={self._initializer.get('random_init')!r} (mechanical_load.py:131-152 knows uniform / normal / gaussian)")
        self._initial_states = self._initializer.get("states", {s: 0.0 for s in self._state_names})

    @property
    def j_total(self):
        return self._j_total

    @property
    def state_names(self):
        return self._state_names

    @property
    def limits(self):
        return self._limits

    @property
    def nominal_values(self):
        return self._nominal_values

    @property
    def initializer(self):
        return self._initializer

    def set_j_rotor(self, j_rotor):
        self._j_total += j_rotor

    def get_state_space(self, omega_range):
        return {"omega": omega_range[0]}, {"omega": omega_range[1]}

    def initial_omega(self):
        return float(self._initial_states.get("omega", 0.0))

    @property
    def random_init(self):
        return self._initializer.get("random_init") in ("uniform", "normal", "gaussian")

    @property
    def gaussian_init(self):
        return self._initializer.get("random_init") in ("normal", "gaussian")

    def gaussian_params(self, lower, upper):
        """(mue, sigma) of the truncated normal (mechanical_load.py:138-141)"""
        rp = self._initializer.get("random_params") or (None, None)
        return float(rp[0] or (upper - lower) / 2 + lower), float(rp[1] or 1)

    def initial_bounds(self, nominal_state, state_low, state_positions):
        """(lower, upper) of the initial omega (mechanical_load.py:118-128)."""
        idx = state_positions["omega"]
        upper = float(nominal_state[idx])
        lower = upper * float(state_low[idx])
        interval = self._initializer.get("interval")
        if interval is not None:
            iv = np.asarray(interval, dtype=float).reshape(-1, 2)
            lower, upper = max(lower, iv[0, 0]), min(upper, iv[0, 1])
        return lower, upper

    def check_initial_state(self, nominal_state, state_low, state_positions):
        """MechanicalLoad.initialize constant branch (mechanical_load.py:151-160)."""
        if self.random_init:
            return
        idx = state_positions["omega"]
        upper = nominal_state[idx]
        lower = upper * state_low[idx]
        if not (lower <= self.initial_omega() <= upper):
            raise Exception("Initialization Value have to be in nominal boundaries")


class ConstantSpeedLoad(MechanicalLoad):
    """reference constant_speed_load.py"""

    KIND = K.LOAD_CONST_SPEED

    def __init__(self, omega_fixed=0, load_initializer=None, **kwargs):
        super().__init__(load_initializer=load_initializer, **kwargs)
        self._omega = omega_fixed or self._initializer["states"]["omega"]
        if omega_fixed != 0:
            self._initializer["states"]["omega"] = omega_fixed
        self._initial_states = self._initializer["states"]

    @property
    def omega_fixed(self):
        return self._omega

    def fill_config(self, cfg):
        cfg.load_kind = self.KIND
        cfg.load_param[K.LP_J_LOAD] = float(self._j_load)


class PolynomialStaticLoad(MechanicalLoad):
    """reference polynomial_static_load.py"""

    KIND = K.LOAD_POLY_STATIC
    _load_parameter = dict(a=0.0, b=0.0, c=0.0, j_load=1e-5)
    tau_decay = 1e-3

    def __init__(self, load_parameter=None, limits=None, load_initializer=None):
        self._load_parameter = update_parameter_dict(self._load_parameter, load_parameter if load_parameter is not None else {})
        super().__init__(j_load=self._load_parameter["j_load"], load_initializer=load_initializer)
        self._limits.update(limits or {})

    @property
    def load_parameter(self):
        return self._load_parameter

    def fill_config(self, cfg):
        cfg.load_kind = self.KIND
        cfg.load_param[K.LP_A] = float(self._load_parameter["a"])
        cfg.load_param[K.LP_B] = float(self._load_parameter["b"])
        cfg.load_param[K.LP_C] = float(self._load_parameter["c"])
        cfg.load_param[K.LP_J_LOAD] = float(self._load_parameter["j_load"])
        cfg.load_param[K.LP_TAU_DECAY] = float(self.tau_decay)


def _unsupported(name, why):
    class _Unsupported(MechanicalLoad):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is not available on the device path: {why}")

    _Unsupported.__name__ = name
    return _Unsupported


class ExternalSpeedLoad(MechanicalLoad):
    """reference external_speed_load.py: the speed follows a user profile f(t) through d omega/dt = (f(t + tau) - omega) / tau.

    A Python callable cannot run inside the kernel, so the profile is TABULATED when the environment is built: f is evaluated at
    every time a fixed-step Euler / RK4 stage can fall on (multiples of tau_system / (2 nsteps)) for `horizon_steps` control steps
    after a reset (t restarts at every reset, as in the reference); past the horizon the last tabulated step repeats.  The scipy
    solvers' adaptive stage times have no table: use EulerSolver / RK4Solver."""

    KIND = K.LOAD_EXT_SPEED
    HAS_JACOBIAN = False

    def __init__(self, speed_profile, load_initializer=None, tau=1e-4, speed_profile_kwargs=None, horizon_steps=100000, **kwargs):
        super().__init__(**kwargs)
        if load_initializer is not None:
            import warnings

            warnings.warn("Given initializer will be overwritten with starting value from speed-profile, to avoid complications at the load "
                          "reset. It is recommended to choose starting value of load by the defined speed-profile.", UserWarning)
        self.speed_profile_kwargs = speed_profile_kwargs or {}
        self._speed_profile = speed_profile
        self._tau = tau
        self._horizon = int(horizon_steps)
        self._omega_initial = float(self._speed_profile(t=0, **self.speed_profile_kwargs))
        self._initializer["states"]["omega"] = self._omega_initial
        self._initial_states = self._initializer["states"]
        self._table = None

    @property
    def omega(self):
        return self._omega_initial

    def fill_config(self, cfg):
        """needs cfg.tau and cfg.solver_nsteps (SCMLSystem.fill_config fills the solver first)"""
        cfg.load_kind = self.KIND
        cfg.load_param[K.LP_J_LOAD] = float(self._j_load)
        cfg.load_param[K.LP_TAU_LOAD] = float(self._tau)
        per = 2 * int(cfg.solver_nsteps)
        dt = float(cfg.tau) / per
        n = per * self._horizon + 2 * per + 1  # two steps of margin: EulerSolver(nsteps > 1) evaluates the profile late (solvers.py:113-119)
        kw = self.speed_profile_kwargs
        try:  # vectorised profiles (numpy expressions) are evaluated in one call
            tab = np.asarray(self._speed_profile(t=np.arange(n) * dt + self._tau, **kw), dtype=np.float64)
            if tab.shape != (n,):
                raise ValueError
        except Exception:
            tab = np.array([float(self._speed_profile(t=j * dt + self._tau, **kw)) for j in range(n)], dtype=np.float64)
        self._table = np.ascontiguousarray(tab)
        cfg.ext_speed_table = self._table.ctypes.data
        cfg.ext_speed_len = n
        cfg._keepalive = self._table  # the C side copies the table in gemb200_create

OrnsteinUhlenbeckLoad = _unsupported("OrnsteinUhlenbeckLoad", "the reference's own constructor raises AttributeError (ornstein_uhlenbeck_load.py:22-27)")
