#!/bin/bash
# ncu --set full captures of one fused-rollout launch and one single-step launch per bench config (tools/one_rollout.py); the reports stay on
# the box (tens of MB each) — what comes back are their raw-metric and source-page CSV exports under gpurun_out/ncu_r02/
mkdir -p gpurun_out/ncu_r02
for cfg in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:"step_kernel|rollout_kernel" -s 1 -c 2 -f -o /tmp/r02_${cfg} python tools/one_rollout.py $cfg 16 1 > gpurun_out/ncu_r02/${cfg}.log 2>&1
  ncu -i /tmp/r02_${cfg}.ncu-rep --page raw --csv > gpurun_out/ncu_r02/${cfg}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_${cfg}.ncu-rep --page source --csv --kernel-id ::regex:rollout_kernel:1 > gpurun_out/ncu_r02/${cfg}_rollout_source.csv 2>/dev/null
  ncu -i /tmp/r02_${cfg}.ncu-rep --page source --csv --kernel-id ::regex:step_kernel:1 > gpurun_out/ncu_r02/${cfg}_step_source.csv 2>/dev/null
  tail -1 gpurun_out/ncu_r02/${cfg}.log
done
du -sh gpurun_out/ncu_r02
