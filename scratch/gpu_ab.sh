for v in "$@"; do
  export GDBGPU_LIB=$PWD/scratch/libgdbgpu_$v.so
  echo "== $v"
  timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -x -q 2>&1 | tail -1
  timeout 600 python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['value']), j['ms_per_step'], j['roofline']['frac'], j['roofline']['kernel_ms_per_step'])"
done
