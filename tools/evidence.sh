#!/bin/bash
# Regenerates the measured evidence kept under profiles/ (run under gpurun; results land in gpurun_out/).
python tools/gpu_check.py --tile-n 256 > gpurun_out/ev_check.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev_bench_ref.json 2>> gpurun_out/ev_bench.err
./tools/tma_bench > gpurun_out/ev_tma_bench.log 2>&1
FVY_DBG=1 python tools/run_layer.py --layers 1,3,6,10,11,27,28,45,53,58 --iters 10 > gpurun_out/ev_timeline.log 2>&1
bash tools/profile.sh > gpurun_out/ev_profile.log 2>&1
tail -3 gpurun_out/ev_check.log; cat gpurun_out/ev_bench_ref.json | cut -c1-300
