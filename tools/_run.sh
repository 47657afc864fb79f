#!/bin/bash
export DPCU_BENCH_C5_TOTAL=67108864
for lw in 0 32; do
  DPCU_BENCH_LINE_WORDS=$([ $lw = 0 ] && echo "" || echo $lw) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t_n2_lw$lw.json 2> gpurun_out/t_n2_lw$lw.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/t_n2_lw$lw.json").read().strip().splitlines()[-1])
c=d["also"]["c5_strong"]
print("line words $lw: c5(64Mi total over 2) step %.4f ms kernel %.4f ms frac %.3f verified %s | headline %.4f ms gather %s" % (c["ms_per_step"], c["avg_launch_ms"], c["frac_of_hbm_peak"], c["bitset_allgather_verified_against_nccl_all_gather"], d["ms_per_step"], d["also"].get("with_bitset_allgather",{}).get("ms_per_step")))
PY
done
