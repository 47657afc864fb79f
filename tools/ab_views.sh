set -e
for v in "2" "3" "4"; do
  touch pipeline_b200/csrc/dpcu_cull.cu
  make -s -j8 -C pipeline_b200/csrc EXTRA="-DDPCU_LINES_MIN_CTAS=$v" > /dev/null 2>&1
  echo "== lines min_ctas=$v"
  python tools/quick_bench.py --views 6 --kernel 4 | sed 's/.*changed=1: //; s/visible.*cull kernel/cull kernel/'
  python tools/quick_bench.py --views 2 --kernel 4 | sed 's/.*changed=1: //; s/visible.*cull kernel/cull kernel/'
done
python tools/quick_bench.py --views 3 --kernel 3 | sed 's/.*changed=1: //; s/visible.*cull kernel/cull kernel/'
python tools/quick_bench.py --views 6 --kernel 2 | sed 's/.*changed=1: //; s/visible.*cull kernel/cull kernel/'
