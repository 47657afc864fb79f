#!/bin/bash
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t_pytest.log 2>&1; tail -3 gpurun_out/t_pytest.log
for w in c2 c3 c4; do python bench.py --workload $w > gpurun_out/t_bench_$w.json 2> gpurun_out/t_bench_$w.err; done
python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
python - <<'PY'
import json
for w in ("c2","c3","c4",""):
    try:
        d=json.loads(open("gpurun_out/t_bench%s.json" % ("_"+w if w else "")).read().strip().splitlines()[-1])
    except Exception as e:
        print(w, "FAILED", e); continue
    r=d["roofline"]; print("%s step %.4f ms value %.2f G e2e %.4f ms kernel %s %.4f frac %.3f share %.2f launches %s" % (d["config"]["workload"], d["ms_per_step"], d["value"]/1e9, d["e2e"]["ms_per_step"], r["kernel"], r["avg_launch_ms"], r["frac"], r["step_share"], d["gpu_launches"]))
    for k,v in d.get("also",{}).items(): print("   also",k,{kk:vv for kk,vv in v.items() if kk in ("ms_per_step","ms","frac_of_hbm_peak","disagreements")})
PY
