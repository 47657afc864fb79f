#!/bin/bash
# round-2 probe B: ncu --set full of the 6-view pair-filter kernel, without / with the L2 prefetch
mkdir -p gpurun_out
for v in _pf0 ""; do
  DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:cullLinesMv -s 3 -c 1 \
    -o gpurun_out/r2b_mv6$v -f python tools/quick_bench.py --views 6 --kernel 7 --iters 3 > gpurun_out/r2b_ncu$v.log 2>&1
  tail -3 gpurun_out/r2b_ncu$v.log
done
ls -la gpurun_out
