#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "conv_tc or tc_dgrad" -p no:cacheprovider -s > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:|worst" gpurun_out/r2r_pytest.log | head -30
