#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r2r_c3_launches.csv python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2r.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2r_c3_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:24]: print(r[4][:70], r[8], r[-1])
PY
