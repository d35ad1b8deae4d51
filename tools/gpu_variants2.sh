#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants2.jsonl
python tools/variant_bench.py variants/libgemb200_base.so variants/libgemb200_b256.so variants/libgemb200_b64.so variants/libgemb200_r56.so variants/libgemb200_stcs.so variants/libgemb200_base.so >> gpurun_out/variants2.jsonl 2>&1
cut -c1-200 gpurun_out/variants2.jsonl
for v in base stcs b256; do GEMB200_LIB=variants/libgemb200_$v.so python bench.py --steps 200 --warmup 20 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('$v', 'bench ms', d['ms_per_step'], 'cold_events', d['cold_events']['ms_per_step'], 'frac', d['roofline']['frac'])"; done
