#!/bin/bash
# Bench lines + launch list of the current build (run under gpurun; results land in gpurun_out/ev8_*).
python bench.py --steps 20 --warmup 5 > gpurun_out/ev8_bench.json 2> gpurun_out/ev8_bench.err
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/ev8_bench_300.json 2>> gpurun_out/ev8_bench.err
python bench.py --size 608 --batch 40 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ev8_bench_608_b40.json 2>> gpurun_out/ev8_bench.err
python bench.py --size 608 --batch 160 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ev8_bench_608_b160.json 2>> gpurun_out/ev8_bench.err
python tools/gpu_check.py --tile-n 256 > gpurun_out/ev8_check.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev8_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ev8_ncu1.log 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for f in bench bench_300 bench_608_b40 bench_608_b160; do cut -c1-160 gpurun_out/ev8_$f.json; echo; done; grep "forward batch" gpurun_out/ev8_check.log; tail -3 gpurun_out/ev8_bench.err
