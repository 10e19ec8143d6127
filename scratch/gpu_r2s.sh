#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2s
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
tail -6 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; tail -c 1200 gpurun_out/${T}_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python scratch/prof_one.py config4 1250 2 > gpurun_out/${T}_launches.log 2>&1
