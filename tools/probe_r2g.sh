#!/bin/bash
# round-2 probe G: full GPU test-suite, then every bench workload at 1 GPU
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
tail -15 gpurun_out/r2g_pytest.log
for w in c4-single c4 c2 c3 c5; do
  timeout 900 python bench.py --workload $w > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err
  echo "bench $w exit $?"
  tail -c 600 gpurun_out/r2g_bench_$w.err
done
python - <<'PY'
import json
for w in ("c4-single","c4","c2","c3","c5"):
    try:
        d=json.loads(open("gpurun_out/r2g_bench_%s.json"%w).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(w, "value %.2f G/s  step %.4f ms  e2e %.4f ms  kernel %s %.4f ms frac %.3f share %.3f launches %d" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], r["kernel"], r["avg_launch_ms"], r["frac"], r["step_share"], d["gpu_launches"]))
        for k,v in d.get("also",{}).items():
            print("   also", k, {kk:vv for kk,vv in v.items() if kk in ("ms_per_step","ms","frac_of_hbm_peak","objects_per_s","disagreements","first_indices","kernel","avg_launch_ms")})
    except Exception as e:
        print(w, "failed", e)
PY
