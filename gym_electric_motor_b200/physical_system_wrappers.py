"""Physical-system wrappers that are fused into the step kernel (reference physical_system_wrappers/*.py).

On the device path a wrapper is a DESCRIPTOR: passing it in `physical_system_wrappers=[...]` switches on the matching
pre-processing inside the kernel (the order of the list has the reference's meaning: later entries wrap earlier ones).

Available: `DqToAbcActionProcessor` (dq_to_abc_action_processor.py), `DeadTimeProcessor` (dead_time_processor.py),
`CurrentSumProcessor` (built into every ShuntDc system, current_sum_processor.py) and the state-vector wrappers
`CosSinProcessor` (cos_sin_processor.py), `FluxObserver` (flux_observer.py) and `StateNoiseProcessor`
(state_noise_processor.py), which run as "state ops" at the end of the step kernel (include/gemb200.h: gemb200_state_op).
"""
import numpy as np

from .spaces import Box


class PhysicalSystemWrapper:
    """reference physical_system_wrapper.py (descriptor)"""


class DqToAbcActionProcessor(PhysicalSystemWrapper):
    """Actions in dq coordinates: a_abc = T32 * q(a_dq, epsilon + angle_advance * tau * omega * p) with angle_advance = 0.5
    (+ the dead time of an inner DeadTimeProcessor), dq_to_abc_action_processor.py:74-95.  PMSM / SynRM: 2 actions; EESM: 3
    (d, q, u_e); SCIM: 2 actions, transformation angle = the FluxObserver state psi_angle (the observer must come first in the list)."""

    def __init__(self, angle_name="epsilon"):
        if angle_name not in ("epsilon", "psi_angle"):
            raise NotImplementedError("transformation angles on the device path: 'epsilon' (PMSM/SynRM/EESM) or the FluxObserver's 'psi_angle' (SCIM)")
        self.angle_name = angle_name

    @classmethod
    def make(cls, motor_type, *args, **kwargs):
        # the reference's registry (dq_to_abc_action_processor.py:96-153): a SynRM uses the "PMSM" entry there, so "SynRM" is refused here too
        assert motor_type in ("PMSM", "EESM", "SCIM", "DFIM"), f"Not supported motor_type {motor_type}."
        if motor_type == "SCIM":  # dq_to_abc_action_processor.py:103-105
            kwargs.setdefault("angle_name", "psi_angle")
        inst = cls(*args, **kwargs)
        inst.dfim = motor_type == "DFIM"  # 4 actions: stator dq + rotor dq (:108-137), needs the FluxObserver's psi_angle as well
        return inst

    dfim = False

    def action_space(self, motor_kind_is_eesm):
        return Box(-1, 1, shape=(4 if self.dfim else (3 if motor_kind_is_eesm else 2),), dtype=np.float64)


class DeadTimeProcessor(PhysicalSystemWrapper):
    """Delays the actions by `steps` control steps (dead_time_processor.py); the queue starts with zero actions after
    every reset (a custom `reset_action` callable is host code and not supported)."""

    def __init__(self, steps=1, reset_action=None):
        if reset_action is not None:
            raise NotImplementedError("custom reset_action callables are host code; the default (zero actions) is built in")
        self._steps = int(steps)
        assert self._steps > 0, f'The number of steps has to be greater than 0. A "{steps}" has been passed.'

    @property
    def dead_time(self):
        return self._steps


class CurrentSumProcessor(PhysicalSystemWrapper):
    """Appends `i_sum`, the sum of the named (normalised) current states (current_sum_processor.py:7-66); its limit / nominal value is the
    maximum or the sum of the source currents' limits.  Built into DcMotorSystem for the shunt motor (the reference's ShuntDc envs wrap
    their system with it); on every other system it runs as a state op of the step kernel (GEMB200_SOP_CURRENT_SUM)."""

    def __init__(self, currents=("i_a", "i_e"), limit="max", physical_system=None):
        assert limit in ["max", "sum"]  # current_sum_processor.py:19
        self._currents = tuple(currents)
        self._limit = limit


class CosSinProcessor(PhysicalSystemWrapper):
    """Appends cos and sin of an angle state (normalised angle * pi) to the state vector, optionally removing the angle
    (cos_sin_processor.py:9-89)."""

    def __init__(self, angle="epsilon", physical_system=None, remove_angle=False):
        self._angle = angle
        self._remove_angle = bool(remove_angle)

    @property
    def angle(self):
        return self._angle


class FluxObserver(PhysicalSystemWrapper):
    """Rotor-flux estimate of an induction motor, appended as `psi_abs`, `psi_angle` (flux_observer.py:9-102)."""

    def __init__(self, current_names=("i_sa", "i_sb", "i_sc"), physical_system=None):
        self._current_names = tuple(current_names)


class StateNoiseProcessor(PhysicalSystemWrapper):
    """Adds i.i.d. noise to the listed states (state_noise_processor.py:4-98).  Distributions on the device path: 'normal'
    (loc, scale), 'uniform' (low, high), 'laplace' (loc, scale); the reference draws blocks of `random_length` samples from numpy,
    the kernel draws one Philox sample per step (same distribution, different stream)."""

    _DISTS = {"normal": ("loc", "scale", 0.0, 1.0), "uniform": ("low", "high", 0.0, 1.0), "laplace": ("loc", "scale", 0.0, 1.0)}

    def __init__(self, states, random_dist="normal", random_kwargs=(), random_length=1000, physical_system=None):
        assert hasattr(np.random.default_rng(), random_dist), (
            f"The numpy random number generator has no distribution {random_dist}."
            "Check https://numpy.org/doc/stable/reference/random/generator.html#distributions for distributions.")
        if random_dist not in self._DISTS:
            raise NotImplementedError(f"random_dist={random_dist!r}: the device path has 'normal', 'uniform' and 'laplace'")
        self._states = states
        self._random_dist = random_dist
        self._random_kwargs = dict(random_kwargs)
        self._random_length = int(random_length)
        k0, k1, d0, d1 = self._DISTS[random_dist]
        unknown = set(self._random_kwargs) - {k0, k1}
        if unknown:
            raise TypeError(f"unexpected random_kwargs {sorted(unknown)} for distribution {random_dist!r}")
        self.params = (float(self._random_kwargs.get(k0, d0)), float(self._random_kwargs.get(k1, d1)))

    @property
    def random_kwargs(self):
        return self._random_kwargs
