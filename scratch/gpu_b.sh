set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_tests.log 2>&1; tail -15 gpurun_out/b_tests.log
timeout 600 python scratch/e2e_probe.py 1250 2 > gpurun_out/b_e2e.log 2>&1; cat gpurun_out/b_e2e.log
