#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2w_smi.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "two_handles or split_schedule" 2>&1 | tail -25 | tee gpurun_out/r2w_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2w_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2w_bench_cfg3_2gpu.json
head -c 700 gpurun_out/r2w_bench_2gpu.json; echo; head -c 700 gpurun_out/r2w_bench_cfg3_2gpu.json
