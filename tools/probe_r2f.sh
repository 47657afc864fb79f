#!/bin/bash
# round-2 probe F: the one-launch grid kernel at 1 Mi objects, GPU-bound step timing
mkdir -p gpurun_out
{
for fl in 0 1 2; do
  for k in 1 8; do
    echo "== 1 Mi objects, kernel $k, flush $fl"
    timeout 300 python tools/quick_bench.py --n 1048576 --views 1 --kernel $k --flush $fl --iters 40
  done
  for v in gp5 gp3 gp0; do
    echo "== 1 Mi objects, kernel 8 ($v), flush $fl"
    DPCU_LIB=$PWD/pipeline_b200/lib/libdpcu_$v.so timeout 300 python tools/quick_bench.py --n 1048576 --views 1 --kernel 8 --flush $fl --iters 40
  done
done
for n in 2097152 4194304; do
  for k in 1 8; do
    echo "== $n objects, kernel $k, flush 2"
    timeout 300 python tools/quick_bench.py --n $n --views 1 --kernel $k --iters 20 --flush 2
  done
done
} > gpurun_out/r2f_bench.log 2>&1
cat gpurun_out/r2f_bench.log
