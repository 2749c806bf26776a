#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a bench step, (2) --set full of representative conv layers.  Run under gpurun.
set -x
python bench.py --steps 2 --warmup 1 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_v5.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/prof_ncu1.log 2>&1
python tools/run_layer.py --layers 28,11,27,3,45,10 --iters 1 > gpurun_out/prof_plain2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 74 -c 12 -o gpurun_out/prof_v5 -f \
    python tools/run_layer.py --layers 28,11,27,3,45,10 --iters 1 > gpurun_out/prof_ncu2.log 2>&1
ncu -i gpurun_out/prof_v5.ncu-rep --page raw --csv > gpurun_out/prof_v5_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
