"""Names of the reference's `visualization` package.  Plotting is out of scope for the device path: the batched env
returns tensors and nothing is drawn.  User scripts written for the reference pass `visualization=MotorDashboard(state_plots=[...])`
to `gem.make`, and agent code looks the class up (`isinstance(v, MotorDashboard)` over `env.visualizations`), so the names resolve to
inert objects: constructing one warns once that nothing will be plotted; every hook is a no-op."""
import warnings

from .core import ElectricMotorVisualization


class _Inert(ElectricMotorVisualization):
    _warned = False

    def __init__(self, *args, update_interval=1000, _quiet=False, **kwargs):
        if not _Inert._warned and not _quiet:
            warnings.warn(f"{type(self).__name__}: gym_electric_motor_b200 does not plot (DESIGN.md §7); the object is accepted and ignored",
                          stacklevel=2)
            _Inert._warned = True
        self.update_interval = update_interval
        self.kwargs = kwargs

    def render(self):
        pass

    def initialize(self):
        pass


class MotorDashboard(_Inert):
    pass


class ConsolePrinter(_Inert):
    pass
