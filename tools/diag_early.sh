#!/bin/bash
# Diagnostic for the early (208^2 / 104^2) layers: role cycle counters + one ncu full capture with source-level stall sampling.
L=${1:-4,6}
FVY_DBG=1 python tools/run_layer.py --layers 1,3,4,6,8,9 --iters 5 2>&1 | grep "fvy dbg" | grep -v timeline > gpurun_out/diag_dbg.log
FVY_CHAIN=0 python tools/run_layer.py --layers $L --iters 1 > gpurun_out/diag_plain.log 2>&1 || exit 1
FVY_CHAIN=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 74 -c 4 -o gpurun_out/diag_early -f \
    python tools/run_layer.py --layers $L --iters 1 > gpurun_out/diag_ncu.log 2>&1
cat gpurun_out/diag_dbg.log
