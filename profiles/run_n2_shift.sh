#!/bin/bash
# 2 GPUs x 12.5 M rows, shared thresholds: tighten milestone sweep
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for sh in 4 6 7 8; do
  ISX_TIGHTEN_SHIFT=$sh timeout 200 $TR --master-port 2963$sh bench.py --gpus 2 --rows 25000000 --steps 5 --warmup 3 --no-cpu-baseline --parity-queries 32 > gpurun_out/r02m_n2_shift$sh.json 2> gpurun_out/r02m_n2_shift$sh.err; echo -n "shift=$sh rc=$? "
  python -c "
import json; d=json.loads(open('gpurun_out/r02m_n2_shift$sh.json').read())
print(round(d['value']), round(d['ms_per_step'],2), d['parity']['bit_exact'], round(d['popc']['candidates_per_query']))"
done
