#!/bin/bash
mkdir -p gpurun_out
{
for v in "" _pf1d1 _pf1d2 _pf1d3 _pf2d2; do
  echo "== lib$v"
  DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$v.so timeout 300 python tools/quick_bench.py --views 6 --kernel 7 --eye 150 0 0 --iters 10
done
} > gpurun_out/r2n.log 2>&1
grep -E "^==|median" gpurun_out/r2n.log | sed -E 's/n=[0-9]+ views=[0-9] kernel=[0-9] ctas=0 fma=0 changed=1: //; s/-> .*visible/| visible/'
