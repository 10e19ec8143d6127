#!/bin/bash
# One GPU call: parity tests, smoke, bench (both arms), config-5 latency, ncu launch list, DRAM traffic of K5, ncu --set full of the top kernels.
set -u
mkdir -p gpurun_out
T=${1:-a}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${T}_gpu.txt 2>&1
nproc >> gpurun_out/${T}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; tail -c 6000 gpurun_out/${T}_bench.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/${T}_bench_ref.log 2>&1; tail -c 1200 gpurun_out/${T}_bench_ref.log
timeout 600 python tools/latency_sweep.py > gpurun_out/${T}_lat.json 2> gpurun_out/${T}_lat.err; cat gpurun_out/${T}_lat.json
timeout 600 python tools/latency_sweep.py --intervals 300 --bcf > gpurun_out/${T}_lat_bcf.json 2>> gpurun_out/${T}_lat.err; cat gpurun_out/${T}_lat_bcf.json
timeout 600 python tools/latency_sweep.py --intervals 300 --records > gpurun_out/${T}_lat_records.json 2>> gpurun_out/${T}_lat.err; cat gpurun_out/${T}_lat_records.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python scratch/prof_one.py config4 1250 2 > gpurun_out/${T}_launches.log 2>&1
tail -2 gpurun_out/${T}_launches.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_remap|k_copy_tile" --csv --log-file gpurun_out/${T}_k5_traffic.csv python scratch/prof_one.py config4 1250 1 > gpurun_out/${T}_k5_traffic.log 2>&1
for k in k_remap_nt k_remap_tile k_copy_tile k_union k_live_walk k_site_scan; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${T}_$k python scratch/prof_one.py config4 313 1 > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls gpurun_out | wc -l
