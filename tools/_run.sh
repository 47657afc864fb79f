#!/bin/bash
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t_pytest.log 2>&1; tail -3 gpurun_out/t_pytest.log
for w in c2 c3 c4; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/t_bench_$w.json 2> gpurun_out/t_bench_$w.err; done
python tools/tree_bench.py 2>&1 | tail -3
python - <<'PY'
import json
for w in ("c2","c3","c4"):
    try:
        d=json.loads(open("gpurun_out/t_bench%s.json" % ("_"+w if w else "")).read().strip().splitlines()[-1])
    except Exception as e:
        print(w, "FAILED", e); continue
    r=d["roofline"]; print("%s step %.4f ms value %.2f G e2e %.4f ms kernel %s %.4f frac %.3f share %.2f launches %s floor %s" % (d["config"]["workload"], d["ms_per_step"], d["value"]/1e9, d["e2e"]["ms_per_step"], r["kernel"], r["avg_launch_ms"], r["frac"], r["step_share"], d["gpu_launches"], r.get("frac_of_cold_read_floor")))
PY
