#!/bin/bash
# Round-2 evidence run (one GPU): bench lines of every workload, the reference arm, ncu launch list of the default bench,
# one `ncu --set full` capture of the dominant kernel of every workload, compute-sanitizer summaries.
# Everything lands in gpurun_out/; tools/ncu_summary.py + a copy step bring the summaries into profiles/ afterwards.
mkdir -p gpurun_out
O=gpurun_out
for w in c4-single c4 c2 c3 c5; do
  timeout 900 python bench.py --workload $w > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; echo "bench $w exit $?"
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "reference arm exit $?"
# launch list of the default command (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
   python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c5 --no-children > $O/r02_launches_bench.log 2>&1; echo "launch list exit $?"
cap() {   # name kernel-regex skip workload objects views
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $O/$1 -f \
     python bench.py --workload $4 --steps 2 --warmup 3 --no-cpu-baseline --no-also > $O/$1.log 2>&1; echo "capture $1 exit $?"
  # summaries here (gpurun brings back at most 64 MiB): the numbers bench.py / DESIGN.md quote, the raw page, the source page
  python tools/ncu_summary.py $O/$1.ncu-rep --workload $4 --objects $5 --views $6 --note "bench.py --workload $4 --steps 2 --warmup 3, 5th launch of the kernel" \
     --out $O/r02_cull_kernel_summary.json > /dev/null
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_ncu_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_ncu_source.csv 2>/dev/null
  if [ "$7" != "keep" ]; then rm -f $O/$1.ncu-rep; fi
}
cap r02_lines1_c4single 'cullLinesKernel' 4 c4-single 67108864 1
cap r02_mv6_c4 'cullLinesMvKernel' 4 c4 67108864 6 keep
cap r02_direct_c2 'cullDirectKernel' 4 c2 1048576 1
cap r02_compact_c2 'compactChangedKernel' 4 c2 1048576 1
cap r02_fused_c3 'cullFusedLeafKernel' 4 c3 16777216 1 keep
cap r02_lines1_c5 'cullLinesKernel' 4 c5 268435456 1
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > $O/r02_sanitizer_$tool.log 2>&1; echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke ok|Error|error" $O/r02_sanitizer_$tool.log | tail -5
done
ls -la $O | tail -40
