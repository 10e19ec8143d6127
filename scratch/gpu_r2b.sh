#!/bin/bash
# ncu --set full of the step's top kernels; the reports stay on the box (too large), summaries + hot source lines come back.
set -u
mkdir -p gpurun_out /tmp/ncu
T=${1:-r2b}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python scratch/prof_one.py config4 1250 2 > gpurun_out/${T}_launches.log 2>&1
for k in k_remap_mix k_union k_live_walk k_site_scan k_verify_alleles k_sig_perm k_inflate; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o /tmp/ncu/${T}_$k python scratch/prof_one.py config4 313 1 > gpurun_out/${T}_ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/${T}_$k.ncu-rep gpurun_out/${T}_${k}_ncu_full.txt
  python tools/ncu_source_hot.py /tmp/ncu/${T}_$k.ncu-rep 30 > gpurun_out/${T}_${k}_source_hot.txt 2>&1
done
bash tools/ncu_traffic.sh > gpurun_out/${T}_traffic.log 2>&1
du -sh gpurun_out
