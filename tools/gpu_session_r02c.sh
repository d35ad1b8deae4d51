#!/bin/bash
# Round 2, third session: full GPU suite (device clock, captured closed loop, single-rank peer stores are new), bench line with the closed-loop
# arm, single-rank peer-gather check, steady-state DRAM traffic (range replay), ncu captures of every bench config.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: (round(v["ms_per_step"] * 1e3, 2), round(v.get("roofline_frac", 0), 3)) for k, v in d.get("other_configs", {}).items()}, d["roofline"]["frac"], d["per_step_launch"]["ms_per_step"], d["e2e"]["value"])
print(d.get("closed_loop_64k"))
P
python tools/peer_gather_check.py --envs 262144 --steps 10 > gpurun_out/peer_check_1rank.log 2>&1; tail -2 gpurun_out/peer_check_1rank.log
for mode in rollout step; do
  ncu --replay-mode range --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/traffic_${mode}_pmsm.csv python tools/traffic_range.py $mode pmsm 8 > gpurun_out/traffic_${mode}_pmsm.log 2>&1
  tail -3 gpurun_out/traffic_${mode}_pmsm.csv; tail -1 gpurun_out/traffic_${mode}_pmsm.log
done
bash tools/gpu_profiles_r02.sh pmsm pmsm_64k fin_sc_pmsm scim eesm
rm -f gpurun_out/ncu_r02/*_step_source.csv
