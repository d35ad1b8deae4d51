"""Voltage-supply descriptors (reference physical_systems/voltage_supplies.py).  Only the ideal supply is on the device
path in this round (all 54 registered envs use it); RC / AC supplies are listed under "next" in DESIGN.md."""


class VoltageSupply:
    """reference voltage_supplies.py:8-57"""

    supply_range = ()
    voltage_len = 1

    def __init__(self, u_nominal):
        self._u_nominal = u_nominal

    @property
    def u_nominal(self):
        return self._u_nominal


class IdealVoltageSupply(VoltageSupply):
    """reference voltage_supplies.py:60-72"""

    def __init__(self, u_nominal=600.0):
        super().__init__(u_nominal)
        self.supply_range = (u_nominal, u_nominal)


def _unsupported(name, where):
    class _Unsupported(VoltageSupply):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} ({where}) is not on the device path yet (SURVEY.md §8f row 3); use IdealVoltageSupply")

    _Unsupported.__name__ = name
    return _Unsupported


RCVoltageSupply = _unsupported("RCVoltageSupply", "voltage_supplies.py:75-123")
AC1PhaseSupply = _unsupported("AC1PhaseSupply", "voltage_supplies.py:126-166")
AC3PhaseSupply = _unsupported("AC3PhaseSupply", "voltage_supplies.py:169-213")
