#!/bin/bash
# Layer-by-layer parity of the conv stack + isolated per-layer times (run under gpurun): gpurun_out/ring1.log
timeout 900 python tools/gpu_check.py --skip-post --tile-n 256 > gpurun_out/ring1.log 2>&1
echo "rc=$?"
grep -c "BAD" gpurun_out/ring1.log
grep "forward batch\|sum of isolated\|head \|Error\|error" gpurun_out/ring1.log | head -20
