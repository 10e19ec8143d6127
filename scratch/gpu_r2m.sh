#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2m
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; tail -c 1500 gpurun_out/${T}_bench.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/${T}_bench_ref.log 2>&1; tail -c 600 gpurun_out/${T}_bench_ref.log
timeout 600 python bench.py --config 2 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_config2.log 2>&1; tail -c 600 gpurun_out/${T}_bench_config2.log
timeout 600 python bench.py --config 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_config3.log 2>&1; tail -c 600 gpurun_out/${T}_bench_config3.log
timeout 600 python tools/latency_sweep.py --intervals 300 --bcf > gpurun_out/${T}_lat_bcf.json 2> gpurun_out/${T}_lat.err; cat gpurun_out/${T}_lat_bcf.json
timeout 600 python tools/latency_sweep.py --intervals 300 --records > gpurun_out/${T}_lat_records.json 2>> gpurun_out/${T}_lat.err; cat gpurun_out/${T}_lat_records.json
timeout 300 python tools/open_bench.py > gpurun_out/${T}_open_bench.json 2>&1
