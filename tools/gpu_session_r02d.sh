#!/bin/bash
# Round 2, fourth session: full GPU suite after the reset-path rework (initial reference values from the walk block, MUFU exp10), bench line,
# ncu raw pages of scim / eesm / pmsm.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: (round(v["ms_per_step"] * 1e3, 2), round(v.get("roofline_frac", 0), 3)) for k, v in d.get("other_configs", {}).items()}, d["roofline"]["frac"], d["per_step_launch"]["ms_per_step"], d["e2e"])
print(d.get("closed_loop_64k"))
P
bash tools/gpu_profiles_r02.sh pmsm scim eesm
rm -f gpurun_out/ncu_r02/*_step_source.csv
