#!/bin/bash
mkdir -p gpurun_out
{
for n in 1572864 2097152 3145728 4194304; do
  for nv in 2 6; do
    for k in 3 7; do
      echo "== $n objects x $nv views, kernel $k, flush 2"
      timeout 300 python tools/quick_bench.py --n $n --views $nv --kernel $k --iters 30 --flush 2 --move 3
    done
  done
done
} > gpurun_out/r2q.log 2>&1
grep -E "^==|median" gpurun_out/r2q.log | sed -E 's/n=[0-9]+ views=[0-9] kernel=[0-9] ctas=0 fma=0 changed=1: //; s/-> .*cull kernel avg/| kernel avg/'
