#!/bin/bash
# One GPU-box session: tests, bench, config sweep, ncu launch list + full capture of the step kernel. Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python tools/config_sweep.py > gpurun_out/config_sweep.jsonl 2> gpurun_out/config_sweep.err; cat gpurun_out/config_sweep.jsonl; tail -3 gpurun_out/config_sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 2 -f -o gpurun_out/prof_step python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
