#!/bin/bash
# round 2: BASELINE configs[3] (training step, global batch 40 @416) at N = 8 with the conv-kernel variants (gpurun --gpus 8)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29911 bench.py --gpus 8 --config train --bucket-mb 256 > gpurun_out/r2m_train_8gpu_mb256_default.json 2> gpurun_out/r2m_train_8gpu_mb256_default.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2m_train_8gpu_mb256_default.json")); x = d["exchange"]
    print("train N=8: %.2f ms/step (cudnn-BN %.2f, bf16 %.2f, tc dgrad %.2f, fvy backward %.2f, fvy all %.2f), all-reduce alone %.3f ms = %.0f GB/s bus, exposed %.3f ms" % (
        d["ms_per_step"], d["baseline_all_cudnn"]["ms_per_step"], d["bf16_autocast"]["ms_per_step"], d["tc_dgrad"]["ms_per_step"], d["fvy_conv_backward"]["ms_per_step"],
        d["fvy_conv_all"]["ms_per_step"], x["allreduce_alone_ms"], x["bus_GBps"], x["exposed_communication_ms"]))
except Exception as e:
    print("train N=8 failed:", e)
PY
tail -3 gpurun_out/r2m_train_8gpu_mb256_default.err
