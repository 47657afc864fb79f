#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "lines_pairs or filter or view or mirror or peer" > gpurun_out/r2l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
{
for v in "" _r0; do
  for e in "0 0 0" "150 0 0"; do
    echo "== lib$v eye $e"
    DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$v.so timeout 300 python tools/quick_bench.py --views 6 --kernel 7 --eye $e --iters 10
  done
  for nv in 2 3 4 8; do
    echo "== lib$v views $nv"
    DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$v.so timeout 300 python tools/quick_bench.py --views $nv --kernel 7 --iters 10
  done
done
} > gpurun_out/r2l.log 2>&1
grep -E "^==|median" gpurun_out/r2l.log | sed -E 's/n=[0-9]+ views=[0-9] kernel=[0-9] ctas=0 fma=0 changed=1: //; s/-> .*visible/| visible/'
