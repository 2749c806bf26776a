#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "wgrad" -p no:cacheprovider -s > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:|worst|Error|error" gpurun_out/r2r_pytest.log | head -40
timeout 300 python tools/conv_bwd_bench.py 2>&1 | tail -10
