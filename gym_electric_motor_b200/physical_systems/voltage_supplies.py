"""Voltage-supply descriptors (reference physical_systems/voltage_supplies.py).  On the device path: the ideal supply (all 54
registered envs use it) and the RC supply (a DC link fed through a resistor, advanced inside the step kernel from the converter's
supply current).  and the single-phase AC supply (phase per env from the Philox stream unless fixed).  The three-phase AC supply
changes the shape of the state vector (three u_sup entries) and is not available."""
import warnings

import numpy as np

from .. import _cabi as K


class VoltageSupply:
    """reference voltage_supplies.py:8-57"""

    supply_range = ()
    voltage_len = 1

    def __init__(self, u_nominal):
        self._u_nominal = u_nominal

    @property
    def u_nominal(self):
        return self._u_nominal


    def fill_config(self, cfg):
        cfg.u_sup = float(self._u_nominal)
        cfg.supply_kind = K.SUPPLY_IDEAL


class IdealVoltageSupply(VoltageSupply):
    """reference voltage_supplies.py:60-72"""

    def __init__(self, u_nominal=600.0):
        super().__init__(u_nominal)
        self.supply_range = (u_nominal, u_nominal)


class RCVoltageSupply(VoltageSupply):
    """reference voltage_supplies.py:75-123: ideal source u_0 = u_nominal behind an RC element,
    d u_sup / dt = (u_0 - u_sup - R i_sup) / (R C), one explicit Euler step per control step."""

    def __init__(self, u_nominal=600.0, supply_parameter=None):
        super().__init__(u_nominal)
        supply_parameter = supply_parameter or {"R": 1, "C": 4e-3}
        assert "R" in supply_parameter.keys(), "Pass key 'R' for Resistance in your dict"
        assert "C" in supply_parameter.keys(), "Pass key 'C' for Capacitance in your dict"
        self.supply_range = (0, u_nominal)
        self._r = supply_parameter["R"]
        self._c = supply_parameter["C"]
        if self._r * self._c < 1e-4:
            warnings.warn("The product of R and C might be too small for the correct calculation of the supply voltage. "
                          "You might want to consider R*C as a time constant.")

    def fill_config(self, cfg):
        cfg.u_sup = float(self._u_nominal)
        cfg.supply_kind = K.SUPPLY_RC
        cfg.supply_param[0], cfg.supply_param[1] = float(self._r), float(self._c)


class AC1PhaseSupply(VoltageSupply):
    """reference voltage_supplies.py:126-166: u_sup(t) = sqrt(2) u_nominal sin(2 pi f t + phi).  Without a 'phase' entry the phase is
    drawn per env at every reset — from the device's Philox stream here, from the unseeded global numpy RNG in the reference."""

    def __init__(self, u_nominal=230, supply_parameter=None):
        super().__init__(u_nominal)
        self._fixed_phi = False
        if supply_parameter is not None:
            assert isinstance(supply_parameter, dict), "supply_parameter should be a dict"
            assert "frequency" in supply_parameter.keys(), "Pass key 'frequency' for frequency f in Hz in your dict"
            supply_parameter = dict(supply_parameter)
            if "phase" in supply_parameter.keys():
                assert 0 <= supply_parameter["phase"] < 2 * np.pi, "The phase angle has to be given in rad in range [0,2*pi)"
                self._fixed_phi = True
            else:
                supply_parameter["phase"] = 0.0
        else:
            supply_parameter = {"frequency": 50, "phase": 0.0}
        self._f = supply_parameter["frequency"]
        self._phi = supply_parameter["phase"]
        self._max_amp = self._u_nominal * np.sqrt(2)
        self.supply_range = [-1 * self._max_amp, self._max_amp]

    def fill_config(self, cfg):
        cfg.u_sup = float(self._u_nominal)
        cfg.supply_kind = K.SUPPLY_AC1
        cfg.supply_param[0], cfg.supply_param[1], cfg.supply_param[2] = float(self._f), float(self._phi), float(self._fixed_phi)


def _unsupported(name, where):
    class _Unsupported(VoltageSupply):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} ({where}) is not on the device path (voltage_len = 3 changes the state vector and the voltage product of physical_systems.py:184, which no registered env or reference test exercises)")

    _Unsupported.__name__ = name
    return _Unsupported


AC3PhaseSupply = _unsupported("AC3PhaseSupply", "voltage_supplies.py:169-213")
