#!/bin/bash
# round-2 probe C: load-pipeline variants of the pair-filter lines kernel (64 Mi objects x 6 views), parity first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "lines_pairs or filter" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
{
for v in "" _a0 _p0 _c5 _c3; do
  echo "== variant libdpcu$v.so, kernel 7"
  for nv in 6; do
    DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$v.so timeout 300 python tools/quick_bench.py --views $nv --kernel 7
  done
done
echo "== default lib, other view counts"
for nv in 2 3 4 5 8; do
  timeout 300 python tools/quick_bench.py --views $nv --kernel 7
done
echo "== skip"
for nv in; do
  DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu_p3.so timeout 300 python tools/quick_bench.py --views $nv --kernel 7
done
} > gpurun_out/r2c_bench.log 2>&1
cat gpurun_out/r2c_bench.log
