#!/bin/bash
# round 2, multi-GPU evidence (gpurun --gpus 8): BASELINE configs[2] (batch 320 @608) at N = 2 / 4 / 8, the headline at N = 8 (e2e with
# uint8 frames), the training step (configs[3]) at N = 8 with exchange variants, and the sharded-identity GPU test across devices.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29700
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "sharded" -p no:cacheprovider > gpurun_out/r2m_sharded.log 2>&1; tail -3 gpurun_out/r2m_sharded.log
for n in 2 4 8; do
  P=$((P+1)); timeout 900 $TR --nproc-per-node $n --master-port $P bench.py --gpus $n --config 608x320 --no-cpu-baseline > gpurun_out/r2m_608x320_${n}gpu.json 2> gpurun_out/r2m_608x320_${n}gpu.err
  cut -c1-200 gpurun_out/r2m_608x320_${n}gpu.json; echo
done
P=$((P+1)); NCCL_DEBUG=INFO timeout 900 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2m_headline_8gpu.json 2> gpurun_out/r2m_headline_8gpu.err
cut -c1-200 gpurun_out/r2m_headline_8gpu.json; echo; grep -c "NCCL INFO" gpurun_out/r2m_headline_8gpu.err; grep -m3 "nranks\|NVLS" gpurun_out/r2m_headline_8gpu.err
for v in "64:" "256:"; do
  mb=${v%%:*}; envv=${v#*:}
  P=$((P+1)); env $envv timeout 900 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --config train --bucket-mb $mb > gpurun_out/r2m_train_8gpu_mb${mb}_${envv:-default}.json 2> gpurun_out/r2m_train_8gpu_mb${mb}_${envv:-default}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2m_train_8gpu_mb${mb}_${envv:-default}.json")); x = d["exchange"]
    print("train N=8 bucket ${mb} MB ${envv}: %.2f ms/step (cudnn-BN %.2f, bf16 %.2f), all-reduce alone %.3f ms = %.0f GB/s bus, exposed %.3f ms" % (d["ms_per_step"], d["baseline_all_cudnn"]["ms_per_step"], d["bf16_autocast"]["ms_per_step"], x["allreduce_alone_ms"], x["bus_GBps"], x["exposed_communication_ms"]))
except Exception as e:
    print("train variant ${mb} ${envv} failed:", e)
PY
done
