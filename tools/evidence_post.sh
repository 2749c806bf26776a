#!/bin/bash
# BASELINE configs[4] (decode / NMS stress): bench line + per-kernel time and DRAM bytes (run under gpurun).
python tools/post_bench.py --steps 20 --warmup 5 > gpurun_out/ev9_post_bench.json 2> gpurun_out/ev9_post_bench.err || { tail -5 gpurun_out/ev9_post_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/ev9_post_kernels.csv \
    python tools/post_bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ev9_ncu.log 2>&1
cat gpurun_out/ev9_post_bench.json; tail -12 gpurun_out/ev9_post_kernels.csv | cut -c1-200
