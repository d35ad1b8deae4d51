"""TEST-ONLY stand-in for the `gymnasium` package (not installed in this image, no network).

It exists solely so that `tests/golden/make_golden.py` can import the UNMODIFIED reference
(`/root/reference/src/gym_electric_motor`) in the build container and record golden trajectories.
It is never on `sys.path` of the product package and must never shadow a real gymnasium install:
`tests/golden/make_golden.py` only adds this directory when `import gymnasium` fails.

Surface = exactly what the reference touches.
"""
import importlib

from . import spaces  # noqa: F401
from .core import Env  # noqa: F401
from .envs.registration import register, registry  # noqa: F401

__version__ = "1.0.0"


def make(env_id, *args, **kwargs):
    spec = registry[env_id]
    mod_name, cls_name = spec["entry_point"].split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    kw = dict(spec["kwargs"])
    kw.update(kwargs)
    kw.pop("disable_env_checker", None)
    return cls(*args, **kw)
