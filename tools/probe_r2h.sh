#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_cuda_parity.py tests/test_full_size.py -x -q -m gpu -k "filter or view or c4 or fma" > gpurun_out/r2h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
{
for e in "0 0 0" "150 0 0" "150 150 150" "600 -300 100"; do
  echo "== eye $e static"
  timeout 300 python tools/quick_bench.py --views 6 --kernel 7 --eye $e --iters 10
done
echo "== moving 3/iter from 0"
timeout 300 python tools/quick_bench.py --views 6 --kernel 7 --move 3 --iters 50
echo "== moving 3/iter from 0, old lines kernel"
timeout 300 python tools/quick_bench.py --views 6 --kernel 4 --move 3 --iters 50
echo "== 4 Mi objects, views kernel, eye 150"
timeout 300 python tools/quick_bench.py --n 4194304 --views 6 --kernel 3 --eye 150 0 0 --iters 10
} > gpurun_out/r2h.log 2>&1
grep -E "^==|median" gpurun_out/r2h.log | sed -E 's/n=[0-9]+ views=6 kernel=[0-9] ctas=0 fma=0 changed=1: //; s/-> .*visible/| visible/'
