"""Reference-generator descriptors (reference core.py:395-509, reference_generators/*.py).  Generation runs in the step
kernel's epilogue with a counter-based Philox stream per (seed, env) — the reference's numpy PCG64 streams cannot be
matched on a device; the distributions are (tests/test_oracle_golden.py::test_wiener_reference_statistics).

On the device: Wiener, Laplace, Sinusoidal, Step, Sawtooth, Triangular, Const, Zero, Multiple and Switched (of those) and
`ExternalReferenceGenerator` (values pushed by the caller each step).  The periodic generators re-derive their sub-episode
parameters from a Philox block addressed by the sub-episode's start step, so they need no extra per-env state.
SwitchedReferenceGenerator keeps (current sub-generator, super-episode end) per env."""
import numpy as np

from . import _cabi as K
from .spaces import Box
from .utils import set_state_array


class ReferenceGenerator:
    """reference core.py:395-509"""

    def __init__(self):
        self.reference_space = None
        self._physical_system = None
        self._referenced_states = None
        self._reference_names = None

    @property
    def referenced_states(self):
        return self._referenced_states

    @property
    def reference_names(self):
        return self._reference_names

    def set_modules(self, physical_system):
        self._physical_system = physical_system

    def slots(self):
        """list of dicts, one per referenced state, consumed by fill_config"""
        raise NotImplementedError

    def fill_config(self, cfg):
        slots = self.slots()
        if len(slots) > K.MAX_REF:
            raise ValueError(f"at most {K.MAX_REF} referenced states")
        cfg.n_ref = len(slots)

        def put(r, s):
            cfg.ref_kind[r] = s["kind"]
            cfg.ref_state[r] = s["state"]
            cfg.ref_value[r] = s.get("value", 0.0)
            cfg.ref_margin_lo[r], cfg.ref_margin_hi[r] = s.get("margin", (-1.0, 1.0))
            cfg.ref_init_lo[r], cfg.ref_init_hi[r] = s.get("init", s.get("margin", (-1.0, 1.0)))
            cfg.ref_sigma_lo[r], cfg.ref_sigma_hi[r] = s.get("sigma", (1e-3, 1e-1))
            cfg.ref_len_lo[r], cfg.ref_len_hi[r] = s.get("length", (500, 2000))
            cfg.ref_amp_lo[r], cfg.ref_amp_hi[r] = s.get("amp", (0.0, 0.0))
            cfg.ref_freq_lo[r], cfg.ref_freq_hi[r] = s.get("freq", (1.0, 1.0))
            cfg.ref_off_lo[r], cfg.ref_off_hi[r] = s.get("off", (0.0, 0.0))

        for r, s in enumerate(slots):
            put(r, s)
        # switched slots: their sub-generators become extra parameter entries behind the output slots (gemb200.h: ref_sw_*)
        nxt = len(slots)
        for r, s in enumerate(slots):
            sw = s.get("switch")
            if not sw or len(sw["subs"]) < 2:
                continue
            m = len(sw["subs"])
            if nxt + m > K.MAX_REF_ENTRIES:
                raise NotImplementedError(f"referenced states + switched sub-generators exceed the {K.MAX_REF_ENTRIES} generator entries of the kernel")
            cfg.ref_sw_count[r], cfg.ref_sw_first[r] = m, nxt
            cfg.ref_sw_len_lo[r], cfg.ref_sw_len_hi[r] = sw["length"]
            acc = 0.0
            for j, sub in enumerate(sw["subs"]):
                put(nxt + j, sub)
                acc += sw["p"][j] / sum(sw["p"])
                cfg.ref_sw_cdf[nxt + j] = min(acc, 1.0) if j < m - 1 else 1.0
            nxt += m

    def close(self):
        pass


class ZeroReferenceGenerator(ReferenceGenerator):
    """reference zero_reference_generator.py"""

    def __init__(self):
        super().__init__()
        self.reference_space = Box(0, 0, (0,), dtype=np.float64)
        self._reference_names = []

    def set_modules(self, physical_system):
        super().set_modules(physical_system)
        self._referenced_states = np.zeros(len(physical_system.state_names), dtype=bool)

    def slots(self):
        return []


class ConstReferenceGenerator(ReferenceGenerator):
    """reference const_reference_generator.py"""

    KIND = K.REF_CONST

    def __init__(self, reference_state="omega", reference_value=0.5, **kwargs):
        super().__init__()
        self._reference_value = reference_value
        self._reference_state = reference_state.lower()
        self.reference_space = Box(np.array([reference_value]), np.array([reference_value]), dtype=np.float64)
        self._reference_names = [self._reference_state]

    def set_modules(self, physical_system):
        super().set_modules(physical_system)
        self._referenced_states = set_state_array({self._reference_state: 1}, physical_system.state_names).astype(bool)

    def slots(self):
        return [dict(kind=self.KIND, state=self._physical_system.state_positions[self._reference_state], value=float(self._reference_value))]


class ExternalReferenceGenerator(ConstReferenceGenerator):
    """Reference values supplied by the caller (`env.set_reference(values)`) — the hook used to inject the reference's
    own reference trajectories in the parity tests, and for user-side generators."""

    KIND = K.REF_EXTERNAL

    def __init__(self, reference_state="omega", initial_value=0.0, **kwargs):
        super().__init__(reference_state=reference_state, reference_value=initial_value)
        self.reference_space = Box(-1, 1, shape=(1,), dtype=np.float64)


class SubepisodedReferenceGenerator(ReferenceGenerator):
    """reference subepisoded_reference_generator.py:9-119"""

    def __init__(self, reference_state="omega", episode_lengths=(500, 2000), limit_margin=None, **kwargs):
        super().__init__()
        self.reference_space = Box(-1, 1, shape=(1,), dtype=np.float64)
        self._limit_margin = limit_margin
        self._reference_state = reference_state.lower()
        self._episode_len_range = episode_lengths
        self._reference_names = [self._reference_state]

    def set_modules(self, physical_system):  # :49-69
        super().set_modules(physical_system)
        ps = physical_system
        self._referenced_states = set_state_array({self._reference_state: 1}, ps.state_names).astype(bool)
        rs = self._referenced_states
        if self._limit_margin is None:
            upper = (ps.nominal_state[rs] / ps.limits[rs])[0] * ps.state_space.high[rs]
            lower = (ps.nominal_state[rs] / ps.limits[rs])[0] * ps.state_space.low[rs]
        elif type(self._limit_margin) in [float, int]:
            upper = self._limit_margin * ps.state_space.high[rs]
            lower = self._limit_margin * ps.state_space.low[rs]
        elif type(self._limit_margin) is tuple:
            lower = self._limit_margin[0] * ps.state_space.low[rs]
            upper = self._limit_margin[1] * ps.state_space.high[rs]
        else:
            raise Exception("Unknown type for the limit margin.")
        self._limit_margin = float(lower[0]), float(upper[0])
        self.reference_space = Box(lower[0], upper[0], shape=(1,), dtype=np.float64)

    def _length_range(self):
        r = self._episode_len_range
        if type(r) in (int, float):
            return int(r), int(r)
        return int(r[0]), int(r[1])


class WienerProcessReferenceGenerator(SubepisodedReferenceGenerator):
    """reference wiener_process_reference_generator.py:7-49"""

    def __init__(self, sigma_range=(1e-3, 1e-1), initial_range=None, **kwargs):
        super().__init__(**kwargs)
        self._initial_range = initial_range
        self._sigma_range = sigma_range

    def set_modules(self, physical_system):
        super().set_modules(physical_system)
        if self._initial_range is None:
            self._initial_range = self._limit_margin

    def slots(self):
        sr = self._sigma_range
        sigma = (float(sr), float(sr)) if type(sr) in (int, float) else (float(sr[0]), float(sr[1]))
        return [dict(kind=K.REF_WIENER, state=self._physical_system.state_positions[self._reference_state], margin=self._limit_margin,
                     init=(float(self._initial_range[0]), float(self._initial_range[1])), sigma=sigma, length=self._length_range())]


class MultipleReferenceGenerator(ReferenceGenerator):
    """reference multiple_reference_generator.py:8-92"""

    def __init__(self, sub_generators, sub_args=None, **kwargs):
        super().__init__()
        self.reference_space = Box(-1, 1, shape=(1,), dtype=np.float64)
        if isinstance(sub_args, dict):
            sub_arguments = [sub_args] * len(sub_generators)
        elif hasattr(sub_args, "__iter__"):
            assert len(sub_args) == len(sub_generators)
            sub_arguments = sub_args
        else:
            sub_arguments = [kwargs] * len(sub_generators)
        self._sub_generators = []
        for sub_generator, sub_arg in zip(sub_generators, sub_arguments):
            if isinstance(sub_generator, str):
                raise Exception
            if isinstance(sub_generator, type):
                sub_generator = sub_generator(**sub_arg)
            self._sub_generators.append(sub_generator)
        self._reference_names = []
        for sub_gen in self._sub_generators:
            self._reference_names += sub_gen.reference_names

    def set_modules(self, physical_system):
        super().set_modules(physical_system)
        for sub in self._sub_generators:
            sub.set_modules(physical_system)
        assert all(sum([sub.referenced_states.astype(int) for sub in self._sub_generators]) < 2), \
            "Some of the passed reference generators share the same reference variable"
        self.reference_space = Box(np.concatenate([s.reference_space.low for s in self._sub_generators]),
                                   np.concatenate([s.reference_space.high for s in self._sub_generators]), dtype=np.float64)
        self._referenced_states = np.sum([s.referenced_states for s in self._sub_generators], dtype=bool, axis=0)

    def slots(self):
        out = []
        for sub in self._sub_generators:
            out += sub.slots()
        return out


class LaplaceProcessReferenceGenerator(SubepisodedReferenceGenerator):
    """reference laplace_process_reference_generator.py: random walk with Laplace(0, sigma) increments, starts at 0."""

    def __init__(self, sigma_range=(1e-3, 1e-1), **kwargs):
        super().__init__(**kwargs)
        self._sigma_range = sigma_range

    def slots(self):
        sr = self._sigma_range
        sigma = (float(sr), float(sr)) if type(sr) in (int, float) else (float(sr[0]), float(sr[1]))
        return [dict(kind=K.REF_LAPLACE, state=self._physical_system.state_positions[self._reference_state], margin=self._limit_margin,
                     sigma=sigma, length=self._length_range())]


class _PeriodicReferenceGenerator(SubepisodedReferenceGenerator):
    """Common part of the sinusoidal / step / sawtooth / triangular generators: per sub-episode a random amplitude, frequency,
    offset and phase (sinusoidal_reference_generator.py:19-62 and siblings)."""

    KIND = None

    def __init__(self, amplitude_range=None, frequency_range=(1, 10), offset_range=None, *_, **kwargs):
        super().__init__(**kwargs)
        self._amplitude_range = amplitude_range or (0, np.inf)
        self._frequency_range = frequency_range
        self._offset_range = offset_range or (-np.inf, np.inf)

    def set_modules(self, physical_system):
        super().set_modules(physical_system)
        self._amplitude_range = np.clip(self._amplitude_range, 0, (self._limit_margin[1] - self._limit_margin[0]) / 2)
        self._offset_range = np.clip(self._offset_range, self._limit_margin[0], self._limit_margin[1])

    @staticmethod
    def _pair(value_range):
        """a number is a fixed value, a pair is a uniform range (SubepisodedReferenceGenerator._get_current_value :102-119)"""
        if np.ndim(value_range) == 0:
            return float(value_range), float(value_range)
        return float(value_range[0]), float(value_range[1])

    def slots(self):
        return [dict(kind=self.KIND, state=self._physical_system.state_positions[self._reference_state], margin=self._limit_margin,
                     amp=self._pair(self._amplitude_range), freq=self._pair(self._frequency_range), off=self._pair(self._offset_range),
                     length=self._length_range())]


class SinusoidalReferenceGenerator(_PeriodicReferenceGenerator):
    """reference sinusoidal_reference_generator.py"""

    KIND = K.REF_SINUS


class StepReferenceGenerator(_PeriodicReferenceGenerator):
    """reference step_reference_generator.py (incl. the roll of the whole sub-episode by int(steps_per_period * phase))"""

    KIND = K.REF_STEP


class SawtoothReferenceGenerator(_PeriodicReferenceGenerator):
    """reference sawtooth_reference_generator.py"""

    KIND = K.REF_SAWTOOTH


class TriangularReferenceGenerator(_PeriodicReferenceGenerator):
    """reference triangle_reference_generator.py"""

    KIND = K.REF_TRIANGULAR


class SwitchedReferenceGenerator(ReferenceGenerator):
    """reference switched_reference_generator.py: switches randomly (probabilities `p`) between sub-generators of the SAME referenced
    state; each one is used for a super-episode of integers(*super_episode_length) steps, and a newly selected sub-generator starts a
    fresh sub-episode from the current reference value.  On the device the sub-generators' parameters occupy additional parameter
    entries of the kernel's generator table (include/gemb200.h: ref_sw_*), so referenced states + extra sub-generators <= 4."""

    def __init__(self, sub_generators, p=None, super_episode_length=(100, 10000)):
        super().__init__()
        self.reference_space = Box(-1, 1, shape=(1,), dtype=np.float64)
        self._sub_generators = list(sub_generators)
        assert len(self._sub_generators) > 0, "No sub generator was passed."
        ref_names = self._sub_generators[0].reference_names
        assert all(sub_gen.reference_names == ref_names for sub_gen in self._sub_generators), \
            "The passed sub generators have different referenced states."
        self._reference_names = ref_names
        self._probabilities = p or [1 / len(sub_generators)] * len(sub_generators)
        if type(super_episode_length) in [float, int]:
            super_episode_length = super_episode_length, super_episode_length + 1
        self._super_episode_length = super_episode_length

    def set_modules(self, physical_system):
        super().set_modules(physical_system)
        for sub_generator in self._sub_generators:
            sub_generator.set_modules(physical_system)
        ref_space_low = np.min([sub.reference_space.low for sub in self._sub_generators], axis=0)
        ref_space_high = np.max([sub.reference_space.high for sub in self._sub_generators], axis=0)
        self.reference_space = Box(ref_space_low, ref_space_high, dtype=np.float64)
        self._referenced_states = self._sub_generators[0].referenced_states
        for sub_generator in self._sub_generators:
            assert np.all(sub_generator.referenced_states == self._referenced_states), "Reference Generators reference different state variables"
            assert sub_generator.reference_space.shape == self.reference_space.shape, "Reference Generators have differently shaped reference spaces"

    def slots(self):
        subs = []
        for sub in self._sub_generators:
            sl = sub.slots()
            if len(sl) != 1 or sl[0]["kind"] == K.REF_EXTERNAL or "switch" in sl[0]:
                raise NotImplementedError("sub-generators of a SwitchedReferenceGenerator must be single-state built-in generators")
            subs.append(sl[0])
        out = dict(subs[0])
        out["switch"] = dict(subs=subs, p=[float(v) for v in self._probabilities],
                             length=(int(self._super_episode_length[0]), int(self._super_episode_length[1])))
        return [out]
