#!/bin/bash
# round-2 probe D: full GPU test-suite, then the one-launch grid kernel at 1 Mi objects (C2) against direct + compaction
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
{
for fl in 0 1 2; do
  for k in 0 1 8 4; do
    echo "== 1 Mi objects, kernel $k, flush $fl"
    timeout 300 python tools/quick_bench.py --n 1048576 --views 1 --kernel $k --flush $fl --iters 30
  done
done
echo "== 1 Mi, no changed list, direct, flush 2"
timeout 300 python tools/quick_bench.py --n 1048576 --views 1 --kernel 1 --flush 2 --iters 30 --changed 0
for n in 4194304 16777216; do
  for k in 1 8 4; do
    echo "== $n objects, kernel $k"
    timeout 300 python tools/quick_bench.py --n $n --views 1 --kernel $k --iters 20
  done
done
for k in 3 8 7; do
  echo "== 1 Mi x 6 views, kernel $k, flush 2"
  timeout 300 python tools/quick_bench.py --n 1048576 --views 6 --kernel $k --flush 2 --iters 30
  echo "== 4 Mi x 6 views, kernel $k"
  timeout 300 python tools/quick_bench.py --n 4194304 --views 6 --kernel $k --iters 30
done
} > gpurun_out/r2d_bench.log 2>&1
cat gpurun_out/r2d_bench.log
