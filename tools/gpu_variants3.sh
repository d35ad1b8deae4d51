#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants3.jsonl
GEMB200_ENV=Cont-CC-EESM-v0 python tools/variant_bench.py variants/libgemb200_f3_m7.so variants/libgemb200_f3_m8.so variants/libgemb200_f3_m6.so variants/libgemb200_f3_m7.so >> gpurun_out/variants3.jsonl 2>&1
GEMB200_ENV=Cont-CC-SCIM-v0 python tools/variant_bench.py variants/libgemb200_f4_m7.so variants/libgemb200_f4_m8.so variants/libgemb200_f4_m6.so variants/libgemb200_f4_m7.so >> gpurun_out/variants3.jsonl 2>&1
cut -c1-140 gpurun_out/variants3.jsonl
