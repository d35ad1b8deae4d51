"""C-ABI checks that need no GPU: the in-tree library loads, exports every symbol include/gemb200.h declares, the ctypes
struct mirror has the C size, config_init/query_dims/validation behave, and the product never imports the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from helpers import ROOT
from gym_electric_motor_b200 import _cabi as K


@pytest.fixture(scope="module")
def lib():
    from gym_electric_motor_b200 import build

    build.build()
    return K.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "gemb200.h")).read()
    declared = sorted(set(re.findall(r"\b(gemb200_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(K.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gemb200_version() == K.ABI_VERSION


def test_struct_mirror_matches_c_layout(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gemb200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(gemb200_config),'
                   ' offsetof(gemb200_config, tau), offsetof(gemb200_config, n_ref), offsetof(gemb200_config, seed));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    size, o_tau, o_nref, o_seed = map(int, subprocess.check_output([str(exe)]).split())
    assert size == C.sizeof(K.GemB200Config)
    assert o_tau == K.GemB200Config.tau.offset and o_nref == K.GemB200Config.n_ref.offset and o_seed == K.GemB200Config.seed.offset


def test_config_init_and_query_dims(lib):
    cfg = K.GemB200Config()
    assert lib.gemb200_config_init(C.byref(cfg)) == 0
    ref = K.new_config()
    assert bytes(cfg) == bytes(ref)  # python-side new_config() == gemb200_config_init
    dims = [C.c_int32() for _ in range(4)]
    for motor, conv, finite, expect in [
        (K.MOTOR_PMSM, (K.CONV_B6, 0), 0, (14, 4, 3)), (K.MOTOR_PMSM, (K.CONV_B6, 0), 1, (14, 4, 1)),
        (K.MOTOR_EESM, (K.CONV_B6, K.CONV_4QC), 0, (16, 5, 4)), (K.MOTOR_EESM, (K.CONV_B6, K.CONV_4QC), 1, (16, 5, 2)),
        (K.MOTOR_SCIM, (K.CONV_B6, 0), 0, (14, 6, 3)), (K.MOTOR_PERMEX_DC, (K.CONV_4QC, 0), 0, (5, 2, 1)),
        (K.MOTOR_SHUNT_DC, (K.CONV_4QC, 0), 0, (7, 3, 1)), (K.MOTOR_EXTEX_DC, (K.CONV_4QC, K.CONV_2QC), 0, (7, 3, 2)),
    ]:
        cfg.motor_kind, cfg.finite = motor, finite
        cfg.converter_kind[0], cfg.converter_kind[1] = conv
        assert lib.gemb200_query_dims(C.byref(cfg), *[C.byref(d) for d in dims]) == 0
        assert tuple(d.value for d in dims[:3]) == expect
    cfg.motor_kind = K.MOTOR_PMSM
    cfg.converter_kind[0], cfg.converter_kind[1] = K.CONV_4QC, 0
    assert lib.gemb200_query_dims(C.byref(cfg), *[C.byref(d) for d in dims]) == K.E_INVALID
    assert b"B6" in lib.gemb200_last_error()


def test_create_validates_before_touching_cuda(lib):
    h = C.c_void_p()
    cfg = K.new_config()
    cfg.struct_size = 8
    assert lib.gemb200_create(C.byref(cfg), C.byref(h)) == K.E_ABI
    cfg = K.new_config()
    cfg.motor_kind, cfg.converter_kind[0] = K.MOTOR_PMSM, K.CONV_B6
    cfg.solver_kind = 7
    assert lib.gemb200_create(C.byref(cfg), C.byref(h)) == K.E_INVALID
    cfg.solver_kind, cfg.interlocking_time = K.SOLVER_RK4, 1.0
    assert lib.gemb200_create(C.byref(cfg), C.byref(h)) == K.E_INVALID
    assert h.value is None


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(K, "_lib", None)
    monkeypatch.setattr(K, "library_path", lambda: str(tmp_path / "libgemb200.so"))
    with pytest.raises(K.GemB200Error, match="no CPU fallback"):
        K.load_library()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gym_electric_motor_b200/ may import, load or reference it."""
    pkg = os.path.join(ROOT, "gym_electric_motor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "gem_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)
    code = "import sys; sys.path.insert(0, %r); import gym_electric_motor_b200 as g; g.make('Cont-CC-PMSM-v0').build_config(); " \
           "assert not any(m.startswith('oracle') for m in sys.modules)" % ROOT
    subprocess.check_call([sys.executable, "-c", code])


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gym_electric_motor_b200 as gem

    env = gem.make("Cont-CC-PMSM-v0", num_envs=4)
    with pytest.raises(K.GemB200Error):
        env.reset()


def _create_rc(lib, cfg):
    h = C.c_void_p()
    rc = lib.gemb200_create(C.byref(cfg), C.byref(h))
    if rc == 0:  # a GPU is present: the handle is real
        lib.gemb200_destroy(h)
    return rc, lib.gemb200_last_error().decode()


def test_every_registered_id_passes_validation(lib):
    """gemb200_create validates before it touches CUDA: without a GPU every registered id must get past validation and fail
    only at the CUDA stage (never E_INVALID / E_ABI); with a GPU it simply succeeds."""
    import gym_electric_motor_b200 as gem
    from gym_electric_motor_b200.envs import env_ids

    for env_id in env_ids():
        cfg = gem.make(env_id, num_envs=8).build_config()
        rc, msg = _create_rc(lib, cfg)
        assert rc in (0, K.E_CUDA), (env_id, rc, msg)


def test_every_golden_config_passes_validation(lib):
    from helpers import config_from_meta, golden_names, load_golden

    for name in golden_names():
        g = load_golden(name)
        for dtype in (K.F64, K.F32):
            # dopri5 goldens: the device runs them as RK4 sub-stepping; the oracle-only solver id must be rejected
            dopri = g["meta"]["case"]["solver"] == "dopri5"
            if dopri:
                rc, msg = _create_rc(lib, config_from_meta(g["meta"], n_envs=4, dtype=dtype))
                assert rc == K.E_INVALID and "solver_kind" in msg
            cfg = config_from_meta(g["meta"], n_envs=4, dtype=dtype, solver="rk4x2" if dopri else None)
            rc, msg = _create_rc(lib, cfg)
            assert rc in (0, K.E_CUDA), (name, rc, msg)
