"""Physical-system wrappers that are fused into the step kernel (reference physical_system_wrappers/*.py).

On the device path a wrapper is a DESCRIPTOR: passing it in `physical_system_wrappers=[...]` switches on the matching
pre-processing inside the kernel (the order of the list has the reference's meaning: later entries wrap earlier ones).

Available: `DqToAbcActionProcessor` (dq_to_abc_action_processor.py), `DeadTimeProcessor` (dead_time_processor.py) and
`CurrentSumProcessor` (built into every ShuntDc system, current_sum_processor.py).  CosSinProcessor, StateNoiseProcessor and
FluxObserver are "next" (SURVEY.md §8f row 1).
"""
import numpy as np

from .spaces import Box


class PhysicalSystemWrapper:
    """reference physical_system_wrapper.py (descriptor)"""


class DqToAbcActionProcessor(PhysicalSystemWrapper):
    """Actions in dq coordinates: a_abc = T32 * q(a_dq, epsilon + angle_advance * tau * omega * p) with angle_advance = 0.5
    (+ the dead time of an inner DeadTimeProcessor), dq_to_abc_action_processor.py:74-95.  PMSM / SynRM: 2 actions; EESM: 3
    (d, q, u_e).  The SCIM variant needs the FluxObserver state 'psi_angle' and is not available yet."""

    def __init__(self, angle_name="epsilon"):
        if angle_name != "epsilon":
            raise NotImplementedError("only the rotor angle 'epsilon' is available as transformation angle (no FluxObserver yet)")
        self.angle_name = angle_name

    @classmethod
    def make(cls, motor_type, *args, **kwargs):
        assert motor_type in ("PMSM", "SynRM", "EESM"), f"Not supported motor_type {motor_type}."
        return cls(*args, **kwargs)

    def action_space(self, motor_kind_is_eesm):
        return Box(-1, 1, shape=(3 if motor_kind_is_eesm else 2,), dtype=np.float64)


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
    """Built into DcMotorSystem for the shunt motor (state `i_sum`); accepted here so that reference-style wrapper lists work."""

    def __init__(self, currents=("i_a", "i_e"), limit="max"):
        if tuple(currents) != ("i_a", "i_e") or limit != "max":
            raise NotImplementedError("only CurrentSumProcessor(('i_a','i_e'), limit='max') — the ShuntDc default — is built in")


def _unsupported(name):
    class _Unsupported(PhysicalSystemWrapper):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is not on the device path yet (SURVEY.md §8f row 1)")

    _Unsupported.__name__ = name
    return _Unsupported


CosSinProcessor = _unsupported("CosSinProcessor")
StateNoiseProcessor = _unsupported("StateNoiseProcessor")
FluxObserver = _unsupported("FluxObserver")
