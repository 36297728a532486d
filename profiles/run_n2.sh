#!/bin/bash
# 2-GPU session: multi-rank parity tests (2 ranks) + cfg3 bench at N=2
set -x
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_store.py -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02f_cfg3_n2.json 2> gpurun_out/r02f_cfg3_n2.err; echo rc=$?; tail -2 gpurun_out/r02f_cfg3_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02f_cfg3_n2.json').read())
print(d['value'], d['ms_per_step'], d['parity'], d['all_gather'])"
