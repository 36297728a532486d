#!/bin/bash
# A/B of library builds on the same box: bench.py (cfg3, 100M and 12.5M rows) per build
for lib in "$@"; do
  for R in 100000000 12500000; do
    echo -n "== $lib rows=$R: "
    ISX_LIB_PATH=$lib python bench.py --rows $R --steps 5 --warmup 3 --no-cpu-baseline --parity-queries 64 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['parity']['bit_exact'], round(d['popc']['candidates_per_query']))"
  done
done
