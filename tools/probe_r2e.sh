#!/bin/bash
# round-2 probe E: ncu of the 1 Mi-object cull (C2): direct kernel vs the one-launch grid kernel
mkdir -p gpurun_out
for k in 1 8; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cull(Direct|Grid)Kernel' -s 3 -c 1 \
    -o gpurun_out/r2e_c2_k$k -f python tools/quick_bench.py --n 1048576 --views 1 --kernel $k --iters 3 > gpurun_out/r2e_ncu_k$k.log 2>&1
  tail -2 gpurun_out/r2e_ncu_k$k.log
done
