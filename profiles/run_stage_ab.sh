#!/bin/bash
# A/B of the staged emission (ISX_STAGE = records per CTA, 0 = direct warp-aggregated emission) on one GPU
for R in 12500000 100000000; do
  for st in 0 1024 256 3072; do
    echo -n "== rows=$R stage=$st: "
    ISX_STAGE=$st python bench.py --rows $R --steps 5 --warmup 3 --no-cpu-baseline --parity-queries 64 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['parity']['bit_exact'], round(d['popc']['candidates_per_query']))"
  done
done
