"""Names of the reference's `visualization` package.  Plotting is out of scope for the device path, but agent code
written against the reference looks the classes up (`isinstance(v, MotorDashboard)` over `env.visualizations`, classic_controllers.py
of the reference's examples), so the names resolve; constructing one says what to use instead."""
from .core import ElectricMotorVisualization


class _NotOnDevice(ElectricMotorVisualization):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__}: matplotlib visualisation is out of scope for gym_electric_motor_b200 (DESIGN.md §7); read the "
                                  "tensors env.step returns, or pass your own ElectricMotorVisualization subclass (its hooks are called)")


class MotorDashboard(_NotOnDevice):
    pass


class ConsolePrinter(_NotOnDevice):
    pass
