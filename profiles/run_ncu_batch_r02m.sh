#!/bin/bash
# ncu --set full of the bulk batch-scan launches (one per word count) of the final round-2 build; reports stay on the box
CMD="python profiles/prof_batch.py --rows 20000000 --queries 10000 --reps 2"
$CMD > gpurun_out/r02m_prof_batch.txt 2>&1 || exit 1
i=0
for spec in "k_scan<.int.2, 26 4" "k_scan<.int.4, 6 3" "k_scan<.int.6, 4 2" "k_scan<.int.8, 2 1"; do
  set -- $spec
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c $3 -f -o /tmp/rep_$i $CMD > gpurun_out/r02m_ncu_$i.log 2>&1
  ncu -i /tmp/rep_$i.ncu-rep --page raw --csv > gpurun_out/r02m_raw_$i.csv 2>/dev/null
  i=$((i+1))
done
python profiles/ncu_summary.py gpurun_out/r02m_ncu_full_batch_scan.csv gpurun_out/r02m_raw_0.csv gpurun_out/r02m_raw_1.csv gpurun_out/r02m_raw_2.csv gpurun_out/r02m_raw_3.csv
cut -d, -f1-4,8-9,11-12 gpurun_out/r02m_ncu_full_batch_scan.csv
