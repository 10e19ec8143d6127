#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2u_topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r2u_bench_n8.log 2>&1
tail -c 6000 gpurun_out/r2u_bench_n8.log
