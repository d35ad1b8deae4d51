"""Build experiment variants of the library into build_variants/ (gpurun-ignored? no: they must travel -> see .gpurunignore).
usage: python tools/build_variants.py [--only-fam=2] name=-DA=1,-DB=2 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_electric_motor_b200 import build as B  # noqa: E402

os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
only_fam = None
specs = []
for a in sys.argv[1:]:
    if a.startswith("--only-fam="):
        only_fam = int(a.split("=")[1])  # e.g. 2 = SYNC: a 6 MB library with the fp32 kernels of one family (fast to ship)
    else:
        specs.append(a)
for spec in specs:
    name, _, defs = spec.partition("=")
    out = os.path.join(ROOT, "variants", f"libgemb200_{name}.so")
    dl = [d for d in defs.split(",") if d]
    if only_fam is not None:
        B.build(force=True, out=out, defines=dl + [f"-DGEMB200_ONLY_FAM={only_fam}"], only={(only_fam, "float")})
    else:
        B.build(force=True, out=out, defines=dl)
    print(out, flush=True)
