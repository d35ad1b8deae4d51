#!/bin/bash
# Round 2, fifth session: GPU suite with the in-kernel clock tick, bench line, ncu launch list of the bench command, compute-sanitizer
# (memcheck over all feature paths incl. the round's new ones, racecheck over a subset).  Outputs under gpurun_out/.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: (round(v["ms_per_step"] * 1e3, 2), round(v.get("roofline_frac", 0), 3)) for k, v in d.get("other_configs", {}).items()}, d["roofline"]["frac"], d["per_step_launch"]["ms_per_step"], d["e2e"]["value"])
print(d.get("closed_loop_64k"))
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-extra > gpurun_out/ncu_launch_bench.log 2>&1; wc -l gpurun_out/launches.csv
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/r02b_sanitize_memcheck.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/r02b_sanitize_memcheck.log
SANITIZE_CASES=0,3,5,16 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/r02b_sanitize_racecheck.log 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/r02b_sanitize_racecheck.log
