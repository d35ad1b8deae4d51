#!/bin/bash
# GPU session for kernel variants: parity tests on the default build, timing of every variant, instruction counts.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python tools/variant_bench.py "" variants/*.so > gpurun_out/variants.jsonl 2>&1
GEMB200_NO_PLAIN=1 python tools/variant_bench.py "" | sed 's/"lib": "default"/"lib": "default_NO_PLAIN"/' >> gpurun_out/variants.jsonl 2>&1
cat gpurun_out/variants.jsonl
for v in variants/libgemb200_fsc.so; do
  GEMB200_LIB=$v python -m pytest tests -m gpu -q -k "(pmsm or synrm or eesm) and f32" 2>&1 | tail -4 > gpurun_out/pytest_$(basename $v .so).log; tail -3 gpurun_out/pytest_$(basename $v .so).log
done
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:step_kernel -s 8 -c 2 --csv --log-file gpurun_out/inst_plain.csv python bench.py --steps 10 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/inst_plain.csv | cut -d, -f5,13- | head -20
