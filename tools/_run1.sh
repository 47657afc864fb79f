python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 2 6; do python -X faulthandler tools/quick_bench.py --views $v --kernel 2 2>&1 | tail -8; done
python tools/quick_bench.py --views 6 --kernel 3 | tail -1
compute-sanitizer --tool memcheck python tools/quick_bench.py --n 300001 --views 3 --kernel 2 --iters 2 2>&1 | tail -2
compute-sanitizer --tool racecheck python tools/quick_bench.py --n 300001 --views 3 --kernel 2 --iters 2 2>&1 | tail -2
