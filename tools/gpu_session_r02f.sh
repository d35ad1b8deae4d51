#!/bin/bash
# Round 2, final session: GPU suite, smoke, bench line, the general-instantiation config (finite + interlocking time), compute-sanitizer.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: (round(v["ms_per_step"] * 1e3, 2), round(v.get("roofline_frac", 0), 3)) for k, v in d.get("other_configs", {}).items()}, d["roofline"]["frac"], d["per_step_launch"]["ms_per_step"], d["e2e"]["value"])
print(d.get("closed_loop_64k"))
P
for c in fin_sc_pmsm_il eesm synrm; do python bench.py --config $c --steps 32 --warmup 4 --no-extra --no-cpu-baseline > gpurun_out/bench_$c.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$c.json')); print('$c', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],3), round(d['per_step_launch']['ms_per_step']*1e3,2))"; done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/r02b_sanitize_memcheck.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/r02b_sanitize_memcheck.log
SANITIZE_CASES=0,3,5,16 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/r02b_sanitize_racecheck.log 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/r02b_sanitize_racecheck.log
