#!/bin/bash
# DRAM traffic of the K5 remap kernel at HEAD, per site of the config-#4 shard (the `roofline.traffic` source of bench.py).
# Run on a GPU box:  bash tools/ncu_traffic.sh   -> gpurun_out/r02_k5_traffic_ncu.csv, profiles/r02_k5_traffic.json
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_remap_mix -s 2 -c 4 --csv \
    --log-file gpurun_out/r02_k5_traffic_ncu.csv python tools/k5_probe.py --no-ab --sites 1250 --steps 1 --variants GDBGPU_K5=mix > gpurun_out/r02_k5_traffic.log 2>&1
python - <<'PY'
import csv, json, re
rows = [r for r in csv.reader(open("gpurun_out/r02_k5_traffic_ncu.csv")) if len(r) > 10]
hdr = rows[0]
i_name, i_metric, i_val, i_unit, i_grid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
per = {}
for r in rows[1:]:
    k = r[0]
    v = float(r[i_val].replace(",", ""))
    u = r[i_unit]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1)   # ncu --csv scales units
    per.setdefault(k, {})[r[i_metric]] = v
log = open("gpurun_out/r02_k5_traffic.log").read()
launches = list(per.values())
tot = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in launches)
out = dict(note="ncu dram__bytes_read.sum + dram__bytes_write.sum of k_remap_mix, config-#4 shard (100k samples x 1250 candidate sites, 8 GiB batches), "
                "launches captured below; bytes per site = bytes of a launch / sites of its batch", launches=launches)
m = re.search(r'"sites_per_step": (\d+)', log)
out["sites_per_step"] = int(m.group(1)) if m else None
json.dump(out, open("gpurun_out/r02_k5_traffic_raw.json", "w"), indent=1)
print(json.dumps(out)[:600])
PY
