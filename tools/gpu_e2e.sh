#!/bin/bash
# e2e (host buffers through gemb200_step_host) vs number of pipeline chunks
for c in 2 4 8 16 32 64; do GEMB200_HOST_CHUNKS=$c python bench.py --steps 60 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('chunks', $c, 'e2e ms', round(d['e2e']['ms_per_step'],4), 'e2e steps/s', '%.3e' % d['e2e']['value'])"; done
