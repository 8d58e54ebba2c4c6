#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2N_bench_${n}gpu.json
  head -c 330 gpurun_out/r2N_bench_${n}gpu.json; echo
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config 3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2N_bench_cfg3_8gpu.json
head -c 330 gpurun_out/r2N_bench_cfg3_8gpu.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2N_bench_cfg4_8gpu.json
head -c 330 gpurun_out/r2N_bench_cfg4_8gpu.json; echo
timeout 200 python -m pytest tests -m gpu -q -k "two_handles" 2>&1 | tail -2
