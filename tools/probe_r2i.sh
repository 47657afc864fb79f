#!/bin/bash
mkdir -p gpurun_out
{
for v in "" _c3; do
  for e in "0 0 0" "150 0 0"; do
    echo "== lib$v eye $e"
    DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$v.so timeout 300 python tools/quick_bench.py --views 6 --kernel 7 --eye $e --iters 10
  done
done
for nv in 2 3 4 5; do
  echo "== lib default views $nv eye 150"
  timeout 300 python tools/quick_bench.py --views $nv --kernel 7 --eye 150 0 0 --iters 10
done
echo "== 8 views"
timeout 300 python tools/quick_bench.py --views 8 --kernel 7 --iters 10
DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu_c3.so timeout 300 python tools/quick_bench.py --views 8 --kernel 7 --iters 10
} > gpurun_out/r2i.log 2>&1
grep -E "^==|median" gpurun_out/r2i.log | sed -E 's/n=[0-9]+ views=[0-9] kernel=[0-9] ctas=0 fma=0 changed=1: //; s/-> .*visible/| visible/'
