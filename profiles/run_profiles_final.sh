#!/bin/bash
# Round-2 evidence: (a) launch list of the default bench command, (b) ncu --set full of the small-batch kernel on the 100M-row
# store (DRAM traffic of a single-query search), (c) ncu --set full of the bulk batch-scan launches (one per word count).
# Reports stay on the box; raw / source pages come back as CSV.
set -x
if [ "$1" != "skip-bench" ]; then
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 32 > gpurun_out/r02g_bench_plain.json 2> gpurun_out/r02g_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02g_launches_bench_n1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 32 > gpurun_out/r02g_bench_under_ncu.log 2>&1
fi
SM="python profiles/prof_batch.py --rows 100000000 --queries 64 --reps 1 --single 1"
$SM || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_scan_small -s 4 -c 12 -f -o /tmp/rep_small $SM > gpurun_out/r02g_ncu_small.log 2>&1
ncu -i /tmp/rep_small.ncu-rep --page raw --csv > gpurun_out/r02g_raw_small.csv 2>/dev/null
CMD="python profiles/prof_batch.py --rows 20000000 --queries 10000 --reps 2"
$CMD || exit 1
i=0
for spec in "k_scan<.int.2, 26 4" "k_scan<.int.4, 6 3" "k_scan<.int.6, 4 2" "k_scan<.int.8, 2 1"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c $3 -f -o /tmp/rep_$i $CMD > gpurun_out/r02g_ncu_$i.log 2>&1
  ncu -i /tmp/rep_$i.ncu-rep --page raw --csv > gpurun_out/r02g_raw_$i.csv 2>/dev/null
  ncu -i /tmp/rep_$i.ncu-rep --page source --csv > gpurun_out/r02g_source_$i.csv 2>/dev/null
  i=$((i+1))
done
ls -la gpurun_out/ | head -30
