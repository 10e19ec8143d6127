#!/bin/bash
set -u
for v in base w4k w2k w4k8; do
  if [ $v = base ]; then unset GDBGPU_LIB; else export GDBGPU_LIB=$PWD/scratch/libgdbgpu_$v.so; fi
  echo "== $v"
  timeout 300 python tools/open_bench.py | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['k_inflate_ms'], j['gzip_open_s'])"
  timeout 300 python -m pytest tests/test_gpu_synthetic.py -x -q -k "compressed or corrupt or inflat" 2>&1 | tail -1
done
