#!/bin/bash
# Focused ncu --set full captures of the batch-regime scan kernels (one bulk launch per word count); raw + source pages
# are exported as CSV on the box (the .ncu-rep files are too large to bring back).
set -x
CMD="python profiles/prof_batch.py --rows 20000000 --queries 10000 --reps 2"
$CMD || exit 1
i=0
for spec in "k_scan<.int.2, 26 4" "k_scan<.int.4, 6 3" "k_scan<.int.6, 4 2" "k_scan<.int.8, 2 1"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c $3 -f -o /tmp/rep_$i $CMD > gpurun_out/r02e_ncu_$i.log 2>&1
  ncu -i /tmp/rep_$i.ncu-rep --page raw --csv > gpurun_out/r02e_raw_$i.csv 2>/dev/null
  ncu -i /tmp/rep_$i.ncu-rep --page source --csv > gpurun_out/r02e_source_$i.csv 2>/dev/null
  i=$((i+1))
done
ls -la gpurun_out/
