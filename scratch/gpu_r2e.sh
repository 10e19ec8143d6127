#!/bin/bash
set -u
mkdir -p gpurun_out
for v in base v1 v2 v3; do
  if [ $v = base ]; then unset GDBGPU_LIB; else export GDBGPU_LIB=$PWD/scratch/libgdbgpu_$v.so; fi
  echo "== $v"
  timeout 600 python tools/k5_probe.py --variants GDBGPU_K5=mix > gpurun_out/r2e_$v.json 2> gpurun_out/r2e_$v.err
  python - <<PY
import json
j=json.load(open("gpurun_out/r2e_$v.json"))
print(j.get("ab")); t=j["timing"]["GDBGPU_K5=mix"]; print(t["sites_per_s"], t["wall_ms_per_step"], t["kernel_ms_per_step"]["k_remap_mix"])
PY
done
unset GDBGPU_LIB
timeout 600 python tools/k5_probe.py --no-ab --variants GDBGPU_K5=mix+GDBGPU_MIX_ONLY=tile,GDBGPU_K5=mix+GDBGPU_MIX_ONLY=nt > gpurun_out/r2e_roles.json 2> gpurun_out/r2e_roles.err
python - <<PY
import json
j=json.load(open("gpurun_out/r2e_roles.json"))
for k,t in j["timing"].items(): print(k, t["wall_ms_per_step"], t["kernel_ms_per_step"]["k_remap_mix"])
PY
