#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "conv_tc or wgrad or fvy_conv_kernels or tc_dgrad" -p no:cacheprovider -s > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:|worst|Error|error" gpurun_out/r2r_pytest.log | head -40
timeout 900 python bench.py --config train > gpurun_out/r2s_train.json 2> gpurun_out/r2s_train.err || tail -5 gpurun_out/r2s_train.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s_train.json"))
print("train N=1: %.2f ms/step | all-cuDNN %.2f | bf16 autocast %.2f | tc dgrad %.2f | fvy backward %.2f | fvy all %.2f" % (d["ms_per_step"], d["baseline_all_cudnn"]["ms_per_step"], d["bf16_autocast"]["ms_per_step"], d["tc_dgrad"]["ms_per_step"], d["fvy_conv_backward"]["ms_per_step"], d["fvy_conv_all"]["ms_per_step"]))
PY
