#!/bin/bash
# sweep of the bootstrap sample size / warm-up ranges on one GPU (cfg3 shape): rows per store given as $1
R=${1:-12500000}
for cfg in "64 2" "128 2" "256 2" "512 2" "1024 2" "256 1" "1024 1" "1024 0" "2048 0"; do
  set -- $cfg
  echo -n "== rows=$R sample=$1 warm=$2: "
  ISX_SAMPLE_BLOCKS=$1 ISX_WARM_RANGES=$2 python bench.py --rows $R --steps 5 --warmup 3 --no-cpu-baseline --parity-queries 64 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['parity']['bit_exact'], round(d['popc']['candidates_per_query']))"
done
