"""ODE-solver descriptors (reference physical_systems/solvers.py).  Integration runs inside the step kernel as a fused
explicit Euler / RK4 sub-stepper (csrc/gemb200_kernels.cuh: integrate<>).

The reference's scipy wrappers (its default is ScipyOdeSolver('dopri5'), rtol 1e-6) have no device twin; they are accepted
for drop-in compatibility and mapped to RK4 with `nsteps=2` sub-steps per tau, which tracks dopri5 to <=1.4e-7 on every
in-scope system (tests/test_gpu_parity.py checks <=2e-6 against dopri5 goldens)."""
from .. import _cabi as K


class OdeSolver:
    """reference solvers.py:4-76 (descriptor only: t / y live on the device)"""

    KIND = K.SOLVER_RK4

    def __init__(self, nsteps=1):
        self._nsteps = int(nsteps)
        if self._nsteps < 1:
            raise ValueError("nsteps must be >= 1")

    @property
    def nsteps(self):
        return self._nsteps

    def fill_config(self, cfg):
        cfg.solver_kind = self.KIND
        cfg.solver_nsteps = self._nsteps


class EulerSolver(OdeSolver):
    """reference solvers.py:79-136"""

    KIND = K.SOLVER_EULER


class RK4Solver(OdeSolver):
    """Classic Runge-Kutta 4 with `nsteps` equal sub-steps per switching segment (no reference twin; see module doc)."""

    KIND = K.SOLVER_RK4


class _ScipyMapped(RK4Solver):
    def __init__(self, *args, **kwargs):
        nsteps = kwargs.pop("nsteps", 2)
        self._scipy_args = (args, kwargs)
        super().__init__(nsteps=nsteps)


class ScipyOdeSolver(_ScipyMapped):
    """reference solvers.py:139-184 — mapped to RK4(nsteps=2)"""


class ScipySolveIvpSolver(_ScipyMapped):
    """reference solvers.py:187-219 — mapped to RK4(nsteps=2)"""


class ScipyOdeIntSolver(_ScipyMapped):
    """reference solvers.py:222-249 — mapped to RK4(nsteps=2)"""
