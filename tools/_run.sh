#!/bin/bash
for lib in "" _nt; do
  echo "== lib$lib"
  export DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$lib.so
  python tools/tree_bench.py 2>&1 | head -1
  python bench.py --workload c3 --no-cpu-baseline --no-also > gpurun_out/t_c3$lib.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/t_c3$lib.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("c3 step %.4f ms e2e %.4f ms kernel %.4f frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], r["avg_launch_ms"], r["frac"]))
PY
done
unset DPCU_LIB
timeout 600 python -m pytest tests -x -q -m gpu -k "tree or fused or leaf or c3" 2>&1 | tail -2
