#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2k_bench_n2.log 2>&1
tail -c 7000 gpurun_out/r2k_bench_n2.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 0 ) > gpurun_out/r2k_bench_ref_n2.log 2>&1
tail -c 1500 gpurun_out/r2k_bench_ref_n2.log
