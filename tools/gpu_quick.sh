mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python tools/config_sweep.py > gpurun_out/config_sweep.jsonl 2> gpurun_out/config_sweep.err; cut -c1-150 gpurun_out/config_sweep.jsonl; tail -3 gpurun_out/config_sweep.err
