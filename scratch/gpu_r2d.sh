#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2d}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
tail -15 gpurun_out/${T}_tests.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-extras > gpurun_out/${T}_bench.log 2>&1; tail -c 3000 gpurun_out/${T}_bench.log
