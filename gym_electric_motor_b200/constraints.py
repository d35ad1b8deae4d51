"""Constraint descriptors + monitor (reference constraints.py, core.py:756-844).  Evaluation is fused into the step
kernel's epilogue; these classes resolve state names into bit masks for the C-ABI."""
import numpy as np

from . import _cabi as K


class Constraint:
    """reference constraints.py:6-29"""

    KIND = None

    def set_modules(self, ps):
        pass

    def mask(self):
        raise NotImplementedError


class LimitConstraint(Constraint):
    """reference constraints.py:32-68: any(|s_i| > 1) over the observed states"""

    KIND = K.CONSTRAINT_LIMIT

    def __init__(self, observed_state_names="all_states"):
        self._observed_state_names = observed_state_names
        self._observed_states = None

    def set_modules(self, ps):
        names = self._observed_state_names
        if names is None:
            names = []
        if "all_states" in names:
            names = ps.state_names
        self._observed_state_names = list(names)
        low = [str(n).lower() for n in self._observed_state_names]
        assert all(n in ps.state_names for n in low), f"A state name in {dict.fromkeys(low).keys()} is invalid."  # utils.set_state_array's wording
        self._observed_states = np.array([n in low for n in ps.state_names], dtype=bool)

    def mask(self):
        return int(sum(1 << i for i, on in enumerate(self._observed_states) if on))


class SquaredConstraint(Constraint):
    """reference constraints.py:71-98: sum(s_i^2) > 1 over the listed (normalised) states"""

    KIND = K.CONSTRAINT_SQUARED

    def __init__(self, states=()):
        self._states = states
        self._state_indices = ()

    def set_modules(self, ps):
        self._state_indices = [ps.state_positions[state] for state in self._states]
        if not np.all(ps.state_space.high[self._state_indices] == 1.0):
            raise NotImplementedError("SquaredConstraint on a non-normalised state space is not supported")

    def mask(self):
        return int(sum(1 << i for i in self._state_indices))


class ConstraintMonitor:
    """reference core.py:756-844 with merge_violations='max' (the only merge the built-in hard constraints need)."""

    def __init__(self, limit_constraints=(), additional_constraints=(), merge_violations="max"):
        self._constraints = list(additional_constraints)
        if len(limit_constraints) > 0:
            self._constraints.append(LimitConstraint(limit_constraints))
        for c in self._constraints:
            if not isinstance(c, Constraint):
                raise TypeError("only LimitConstraint / SquaredConstraint instances can be fused into the device epilogue; "
                                "callable constraints are host code (INTEGRATION.md)")
        if merge_violations not in ("max", "product"):
            raise NotImplementedError("callable merge_violations is host code and not supported")
        # for violation degrees in {0, 1} 'max' and 'product' coincide (1 - prod(1 - v))
        self._merge = merge_violations
        if len(self._constraints) > K.MAX_CONSTRAINTS:
            raise ValueError(f"at most {K.MAX_CONSTRAINTS} constraints")

    @property
    def constraints(self):
        return self._constraints

    def set_modules(self, ps):
        for c in self._constraints:
            c.set_modules(ps)

    def fill_config(self, cfg):
        cfg.n_constraints = len(self._constraints)
        for i, c in enumerate(self._constraints):
            cfg.constraint_kind[i] = c.KIND
            cfg.constraint_mask[i] = c.mask()
