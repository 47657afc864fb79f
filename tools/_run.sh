#!/bin/bash
python tools/stream_floor.py
timeout 900 python -m pytest tests -x -q -m gpu -k "fused or tree or c3 or leaf" 2>&1 | tail -2
(time python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open("gpurun_out/t_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("%s step %.4f ms value %.2f G e2e %.4f ms kernel %s %.4f frac %.3f share %.2f launches %s" % (d["config"]["workload"], d["ms_per_step"], d["value"]/1e9, d["e2e"]["ms_per_step"], r["kernel"], r["avg_launch_ms"], r["frac"], r["step_share"], d["gpu_launches"]))
for k,v in d.get("also",{}).items(): print("   also",k,json.dumps({kk:vv for kk,vv in v.items() if kk in ("ms_per_step","ms","frac_of_hbm_peak","disagreements","e2e_ms_per_step","roofline","failed")})[:400])
PY
