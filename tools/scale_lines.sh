#!/bin/bash
# N-GPU bench lines: bash tools/scale_lines.sh N (default workload with also.c5_strong, and the sharded C3 tree)
mkdir -p gpurun_out
N=${1:-2}
for w in c4-single c3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 20 --warmup 3 > gpurun_out/r02_scale_n${N}_$w.json 2> gpurun_out/r02_scale_n${N}_$w.err
  echo "bench $w N=$N exit $?"
  tail -c 800 gpurun_out/r02_scale_n${N}_$w.err
done
python - <<PY
import json
for w in ("c4-single","c3"):
    try:
        d=json.loads(open("gpurun_out/r02_scale_n${N}_%s.json"%w).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(w, "N=%d value %.2f G/s  step %.4f ms  e2e %.4f ms  kernel %s %.4f ms frac %.3f launches %d" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], r["kernel"], r["avg_launch_ms"], r["frac"], d["gpu_launches"]))
        for k,v in d.get("also",{}).items():
            print("   also", k, {kk:vv for kk,vv in v.items() if kk in ("ms_per_step","ms","frac_of_hbm_peak","objects_per_s","kernel","avg_launch_ms","verified_against_nccl_all_gather","bitset_allgather_verified_against_nccl_all_gather")})
    except Exception as e:
        print(w, "failed", e)
PY
