#!/bin/bash
set -u
for v in base va vc ve; do
  if [ $v = base ]; then unset GDBGPU_LIB; else export GDBGPU_LIB=$PWD/scratch/libgdbgpu_$v.so; fi
  echo "== $v"
  timeout 600 python tools/k5_probe.py --variants GDBGPU_K5=mix > /tmp/r2v_$v.json 2> /tmp/r2v_$v.err
  python - <<PY
import json
j=json.load(open("/tmp/r2v_$v.json"))
print(j.get("ab_identical"), [v["sha256"] for v in j["ab"].values()]); t=j["timing"]["GDBGPU_K5=mix"]; print(t["sites_per_s"], t["wall_ms_per_step"], t["kernel_ms_per_step"]["k_remap_mix"])
PY
done
