#!/bin/bash
# full GPU check of the committed state: test-suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/full_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/full_pytest.log
tail -3 gpurun_out/full_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/full_bench_ref.json 2>/dev/null; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/full_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.2f G/s step %.4f ms e2e %.2f G/s (%.4f ms) kernel %s frac %.3f traffic %s launches %d clocks %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], r["kernel"], r["frac"], r["traffic"], d["gpu_launches"], d["clocks"]))
for k,v in d["also"].items(): print("  also",k,{kk:vv for kk,vv in v.items() if kk in ("ms_per_step","ms","frac_of_hbm_peak","disagreements")})
print("cpu_baseline", d["cpu_baseline"]["value"]/1e6, "M/s", d["cpu_baseline"]["cores"], "core")
r=json.loads(open("gpurun_out/full_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm %.1f M objects/s on %d threads; e2e ratio %.1fx" % (r["value"]/1e6, r["cpu_baseline"]["cores"], d["e2e"]["value"]/r["value"]))
PY
