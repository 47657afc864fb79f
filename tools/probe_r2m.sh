#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cullLinesMv -s 3 -c 1 -o gpurun_out/r2m_ring -f python tools/quick_bench.py --views 6 --kernel 7 --iters 3 > gpurun_out/r2m_ncu.log 2>&1
ncu -i gpurun_out/r2m_ring.ncu-rep --page raw --csv > gpurun_out/r2m_ring_raw.csv 2>/dev/null
ncu -i gpurun_out/r2m_ring.ncu-rep --page source --csv > gpurun_out/r2m_ring_source.csv 2>/dev/null
rm -f gpurun_out/r2m_ring.ncu-rep
