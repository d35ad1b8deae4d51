"""Power-electronic converter descriptors (reference physical_systems/converters.py).  The conversion itself runs in the
step kernel (csrc/gemb200_kernels.cuh: c2qc / cont_qc / f2qc_leg / f2qc_out); these classes carry the kind, the action /
voltage / current spaces and the interlocking time, with the reference's class names and constructor kwargs."""
import numpy as np

from .. import _cabi as K
from ..spaces import Box, Discrete, MultiDiscrete


class PowerElectronicConverter:
    """reference converters.py:5-111"""

    voltages = None
    currents = None
    action_space = None
    KIND = K.CONV_NONE
    FINITE = False

    def __init__(self, tau=1e-4, interlocking_time=0.0):
        self._tau = float(tau)
        self._interlocking_time = float(interlocking_time)

    @property
    def tau(self):
        return self._tau

    @tau.setter
    def tau(self, value):
        self._tau = float(value)

    @property
    def interlocking_time(self):
        return self._interlocking_time

    def slots(self):
        """[(kind), ...] for the C-ABI converter_kind[] (one or two slots)."""
        return [self.KIND]


class ContDynamicallyAveragedConverter(PowerElectronicConverter):
    """reference converters.py:130-184"""

    def __init__(self, tau=1e-4, **kwargs):
        super().__init__(tau=tau, **kwargs)


class FiniteConverter(PowerElectronicConverter):
    """reference converters.py:187-215"""

    FINITE = True

    def __init__(self, tau=1e-5, **kwargs):
        super().__init__(tau=tau, **kwargs)


class ContOneQuadrantConverter(ContDynamicallyAveragedConverter):
    KIND = K.CONV_1QC
    voltages = Box(0, 1, shape=(1,), dtype=np.float64)
    currents = Box(0, 1, shape=(1,), dtype=np.float64)
    action_space = Box(0, 1, shape=(1,), dtype=np.float64)


class ContTwoQuadrantConverter(ContDynamicallyAveragedConverter):
    KIND = K.CONV_2QC
    voltages = Box(0, 1, shape=(1,), dtype=np.float64)
    currents = Box(-1, 1, shape=(1,), dtype=np.float64)
    action_space = Box(0, 1, shape=(1,), dtype=np.float64)


class ContFourQuadrantConverter(ContDynamicallyAveragedConverter):
    KIND = K.CONV_4QC
    voltages = Box(-1, 1, shape=(1,), dtype=np.float64)
    currents = Box(-1, 1, shape=(1,), dtype=np.float64)
    action_space = Box(-1, 1, shape=(1,), dtype=np.float64)


class ContB6BridgeConverter(ContDynamicallyAveragedConverter):
    KIND = K.CONV_B6
    action_space = Box(-1, 1, shape=(3,), dtype=np.float64)
    voltages = Box(-1, 1, shape=(3,), dtype=np.float64)
    currents = Box(-1, 1, shape=(3,), dtype=np.float64)


class FiniteOneQuadrantConverter(FiniteConverter):
    KIND = K.CONV_1QC
    voltages = Box(0, 1, shape=(1,), dtype=np.float64)
    currents = Box(0, 1, shape=(1,), dtype=np.float64)
    action_space = Discrete(2)


class FiniteTwoQuadrantConverter(FiniteConverter):
    KIND = K.CONV_2QC
    voltages = Box(0, 1, shape=(1,), dtype=np.float64)
    currents = Box(-1, 1, shape=(1,), dtype=np.float64)
    action_space = Discrete(3)


class FiniteFourQuadrantConverter(FiniteConverter):
    KIND = K.CONV_4QC
    voltages = Box(-1, 1, shape=(1,), dtype=np.float64)
    currents = Box(-1, 1, shape=(1,), dtype=np.float64)
    action_space = Discrete(4)


class FiniteB6BridgeConverter(FiniteConverter):
    KIND = K.CONV_B6
    action_space = Discrete(8)
    voltages = Box(-1, 1, shape=(3,), dtype=np.float64)
    currents = Box(-1, 1, shape=(3,), dtype=np.float64)


class _MultiMixin:
    def _init_multi(self, subconverters, kwargs):
        self._sub_converters = []
        for sub in subconverters:
            assert not isinstance(sub, str)
            if isinstance(sub, type):
                sub = sub(**kwargs)
            if isinstance(sub, (ContMultiConverter, FiniteMultiConverter)):
                raise TypeError("sub-converters must be elementary converters (reference converters.py:500-501)")
            if sub.FINITE != self.FINITE:
                raise TypeError("cannot mix finite and continuous sub-converters")
            self._sub_converters.append(sub)
        if len(self._sub_converters) > 2:
            raise NotImplementedError("at most two sub-converters are supported on the device")
        # `_interlocking_time` stays the multi converter's own (unused) kwarg like in the reference (converters.py:629-636, DESIGN.md finding 5);
        # what the half bridges actually use are the sub-converters' values — one per converter slot of the kernel (gemb200.h: interlocking_time,
        # interlocking_time1)
        self._sub_interlocking_times = [float(s.interlocking_time) for s in self._sub_converters] or [float(self._interlocking_time)]
        self._sub_interlocking_time = self._sub_interlocking_times[0]
        self.currents = Box(np.concatenate([s.currents.low for s in self._sub_converters]),
                            np.concatenate([s.currents.high for s in self._sub_converters]), dtype=np.float64)
        self.voltages = Box(np.concatenate([s.voltages.low for s in self._sub_converters]),
                            np.concatenate([s.voltages.high for s in self._sub_converters]), dtype=np.float64)

    def sub_converters(self):
        return self._sub_converters

    @property
    def interlocking_time(self):
        return self._sub_interlocking_time

    def interlocking_times(self):
        """interlocking time per converter slot (sub-converter)"""
        return list(self._sub_interlocking_times)

    def slots(self):
        return [s.KIND for s in self._sub_converters]


class ContMultiConverter(_MultiMixin, ContDynamicallyAveragedConverter):
    """reference converters.py:615-740"""

    def __init__(self, subconverters, **kwargs):
        ContDynamicallyAveragedConverter.__init__(self, **kwargs)
        self._init_multi(subconverters, kwargs)
        self.action_space = Box(np.concatenate([s.action_space.low for s in self._sub_converters]),
                                np.concatenate([s.action_space.high for s in self._sub_converters]), dtype=np.float64)

    @PowerElectronicConverter.tau.setter
    def tau(self, value):
        self._tau = float(value)
        for s in self._sub_converters:
            s.tau = value


class FiniteMultiConverter(_MultiMixin, FiniteConverter):
    """reference converters.py:498-612"""

    def __init__(self, subconverters, **kwargs):
        FiniteConverter.__init__(self, **kwargs)
        self._init_multi(subconverters, kwargs)
        self.action_space = MultiDiscrete([s.action_space.n for s in self._sub_converters])

    @PowerElectronicConverter.tau.setter
    def tau(self, value):
        self._tau = float(value)
        for s in self._sub_converters:
            s.tau = value
