#!/bin/bash
# 2-GPU session: multi-rank parity tests (2 ranks, shared thresholds) + cfg3 bench at N=2, staged emission off / on
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for st in 0 1024; do
  ISX_STAGE=$st timeout 300 $TR --master-port 2962$((st % 7)) bench.py --gpus 2 --rows 25000000 --steps 5 --warmup 3 --no-cpu-baseline --parity-queries 64 > gpurun_out/r02l_n2_stage$st.json 2> gpurun_out/r02l_n2_stage$st.err; echo "stage=$st rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r02l_n2_stage$st.json').read())
print(round(d['value']), round(d['ms_per_step'],2), d['parity']['bit_exact'], d['parity'].get('shared_thresholds'), round(d['popc']['candidates_per_query']))"
done
