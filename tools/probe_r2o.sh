#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "grid or kernel_choice" > gpurun_out/r2o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
{
for n in 1048576 2097152 4194304 16777216; do
  for k in 1 8; do
    echo "== $n objects, kernel $k, flush 2"
    timeout 300 python tools/quick_bench.py --n $n --views 1 --kernel $k --iters 30 --flush 2
  done
done
echo "== 1 Mi, kernel 8, no flush"
timeout 300 python tools/quick_bench.py --n 1048576 --views 1 --kernel 8 --iters 30
echo "== 1 Mi, kernel 1, no flush"
timeout 300 python tools/quick_bench.py --n 1048576 --views 1 --kernel 1 --iters 30
for k in 3 8; do
  echo "== 1 Mi x 6 views, kernel $k, flush 2"
  timeout 300 python tools/quick_bench.py --n 1048576 --views 6 --kernel $k --flush 2 --iters 30
  echo "== 4 Mi x 6 views, kernel $k, flush 2"
  timeout 300 python tools/quick_bench.py --n 4194304 --views 6 --kernel $k --iters 30 --flush 2
done
} > gpurun_out/r2o.log 2>&1
grep -E "^==|median" gpurun_out/r2o.log | sed -E 's/n=[0-9]+ views=[0-9] kernel=[0-9] ctas=0 fma=0 changed=1: //; s/-> .*cull kernel avg/| kernel avg/'
