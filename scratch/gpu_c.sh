set -u
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/c_bench.log 2>&1; tail -c 5000 gpurun_out/c_bench.log
timeout 600 python tools/latency_sweep.py > gpurun_out/c_lat.json 2> gpurun_out/c_lat.err; cat gpurun_out/c_lat.json; tail -3 gpurun_out/c_lat.err
timeout 600 python tools/latency_sweep.py --intervals 300 --bcf > gpurun_out/c_lat_bcf.json 2>> gpurun_out/c_lat.err; cat gpurun_out/c_lat_bcf.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_remap|k_copy_tile" --csv --log-file gpurun_out/c_k5_traffic.csv python scratch/prof_one.py config4 1250 1 > gpurun_out/c_k5_traffic.log 2>&1
tail -2 gpurun_out/c_k5_traffic.log
