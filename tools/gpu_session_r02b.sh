#!/bin/bash
# Round 2, second half: full GPU suite, register-cap variants of the rollout kernel, bench line, ncu captures (raw + source pages) of the
# PMSM and EESM rollout / step kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python tools/rollout_variant_bench.py "$@" > gpurun_out/rollout_variants.jsonl 2>&1; cat gpurun_out/rollout_variants.jsonl
python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: (round(v["ms_per_step"] * 1e3, 2), round(v.get("roofline_frac", 0), 3)) for k, v in d.get("other_configs", {}).items()}, d["roofline"]["frac"], d["per_step_launch"]["ms_per_step"], d["e2e"]["value"])
P
bash tools/gpu_profiles_r02.sh pmsm eesm
