#!/bin/bash
# What the driver runs at round end, in one session: GPU suite, smoke, the reference arm and the product arm of bench.py with the driver's flags.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_reference.json) 2>&1 | tail -3; cut -c1-300 gpurun_out/bench_final_reference.json
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err) 2>&1 | tail -3; cut -c1-420 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
