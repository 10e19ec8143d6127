#!/bin/bash
# Round-2 HEAD measurement: parity tests, smoke, bench, launch list, K5 traffic, ncu --set full of the top kernels.
set -u
mkdir -p gpurun_out
T=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${T}_gpu.txt 2>&1
nproc >> gpurun_out/${T}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/${T}_tests.log 2>&1
tail -25 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; tail -c 8000 gpurun_out/${T}_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python scratch/prof_one.py config4 1250 2 > gpurun_out/${T}_launches.log 2>&1
tail -2 gpurun_out/${T}_launches.log
timeout 600 bash tools/ncu_traffic.sh > gpurun_out/${T}_traffic.log 2>&1; tail -3 gpurun_out/${T}_traffic.log
for k in k_remap_mix k_union k_live_walk k_site_scan k_verify_alleles; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${T}_$k python scratch/prof_one.py config4 313 1 > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls gpurun_out | wc -l
