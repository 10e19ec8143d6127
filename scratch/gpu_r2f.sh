#!/bin/bash
set -u
mkdir -p gpurun_out /tmp/ncu
export GDBGPU_K5=mix GDBGPU_MIX_ONLY=tile
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_remap_mix -s 2 -c 1 -f -o /tmp/ncu/tileonly python tools/k5_probe.py --no-ab --steps 1 --variants GDBGPU_K5=mix+GDBGPU_MIX_ONLY=tile > gpurun_out/r2f_ncu.log 2>&1
ncu -i /tmp/ncu/tileonly.ncu-rep --page raw --csv > /tmp/ncu/raw.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/ncu/raw.csv')))
hdr,units,r=rows[0],rows[1],rows[2]
out=open('gpurun_out/r2f_tileonly_raw.txt','w')
for h,u,v in zip(hdr,units,r):
    if any(k in h for k in ('stalled','throughput','issue_active','warps_active','dram__bytes','time_duration','inst_executed.sum','lsu','hit_rate','bank_conflict','occupancy','registers','shared_mem')):
        out.write("%-100s %s %s\n"%(h,v,u))
PY
python tools/ncu_source_hot.py /tmp/ncu/tileonly.ncu-rep 40 > gpurun_out/r2f_tileonly_source_hot.txt 2>&1
