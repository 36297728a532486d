#!/bin/bash
# One 8-GPU session: multi-rank parity tests, then BASELINE configs 3, 4, 5 at full size (profiles/r02_*_n8.json).
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc; free -g | head -2
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02i_cfg3_n8.json 2> gpurun_out/r02i_cfg3_n8.err; echo rc=$?; tail -3 gpurun_out/r02i_cfg3_n8.err
timeout 400 $TR --master-port 29612 bench.py --config cfg4 --gpus 8 --steps 5 --warmup 2 > gpurun_out/r02i_cfg4_n8.json 2> gpurun_out/r02i_cfg4_n8.err; echo rc=$?; tail -3 gpurun_out/r02i_cfg4_n8.err
timeout 600 $TR --master-port 29613 bench.py --config cfg5 --gpus 8 --steps 1 --warmup 1 > gpurun_out/r02i_cfg5_n8.json 2> gpurun_out/r02i_cfg5_n8.err; echo rc=$?; tail -3 gpurun_out/r02i_cfg5_n8.err
