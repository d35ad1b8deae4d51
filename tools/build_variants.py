"""Build experiment variants of the library into build_variants/ (gpurun-ignored? no: they must travel -> see .gpurunignore).
usage: python tools/build_variants.py name=-DA=1,-DB=2 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_electric_motor_b200 import build as B  # noqa: E402

os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    out = os.path.join(ROOT, "variants", f"libgemb200_{name}.so")
    B.build(force=True, out=out, defines=[d for d in defs.split(",") if d])
    print(out, flush=True)
