#!/bin/bash
# round-2 probe A: parity of the pair-filter lines kernel, then timing of its tuning variants (64 Mi objects x 6 views)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "lines_pairs or filter" > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
{
echo "== baseline old lines kernel (4), 6 views"
timeout 300 python tools/quick_bench.py --views 6 --kernel 4
for v in "" _pf0 _pf1 _pf1d4 _pf2d4 _c5 _c3; do
  echo "== variant libdpcu$v.so, kernel 7"
  for nv in 6; do
    DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$v.so timeout 300 python tools/quick_bench.py --views $nv --kernel 7
  done
done
echo "== default lib, other view counts"
for nv in 2 3 4 8; do
  timeout 300 python tools/quick_bench.py --views $nv --kernel 7
  timeout 300 python tools/quick_bench.py --views $nv --kernel 4
done
} > gpurun_out/r2a_bench.log 2>&1
cat gpurun_out/r2a_bench.log
