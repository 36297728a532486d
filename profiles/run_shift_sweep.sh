#!/bin/bash
# sweep of the tighten milestone (2^shift candidates of a query between two threshold re-derivations) with the staged emission
for R in 12500000 100000000; do
  for sh in 2 3 4 5 6; do
    [ $R = 100000000 ] && [ $sh = 2 -o $sh = 6 ] && continue
    echo -n "== rows=$R shift=$sh: "
    ISX_TIGHTEN_SHIFT=$sh python bench.py --rows $R --steps 5 --warmup 3 --no-cpu-baseline --parity-queries 32 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['parity']['bit_exact'], round(d['popc']['candidates_per_query']))"
  done
done
