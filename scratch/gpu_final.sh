#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2z
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/${T}_tests.log 2>&1
tail -14 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; tail -c 800 gpurun_out/${T}_bench.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/${T}_bench_ref.log 2>&1; tail -c 300 gpurun_out/${T}_bench_ref.log
