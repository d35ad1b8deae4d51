#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/pf.jsonl
for d in 0 37888 75776 151552 303104 606208; do
  GEMB200_PF_DIST=$d python tools/variant_bench.py "" | sed "s/\"lib\": \"default\"/\"pf_dist\": $d/" >> gpurun_out/pf.jsonl
done
cat gpurun_out/pf.jsonl
for d in 0 151552; do GEMB200_PF_DIST=$d python bench.py --steps 200 --warmup 20 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('pf', $d, 'bench ms', d['ms_per_step'], 'cold_events', d['cold_events']['ms_per_step'], 'frac', d['roofline']['frac'])"; done
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
