#!/bin/bash
# round 2, multi-GPU evidence of the final inference build (gpurun --gpus 8): sharded identity, BASELINE configs[2] at N = 2 / 4 / 8,
# the headline at N = 2 / 4 / 8.  (The training lines of tools/r2_run_multi.sh are unchanged by the inference work.)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29800
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "sharded" -p no:cacheprovider > gpurun_out/r2m_sharded.log 2>&1; tail -3 gpurun_out/r2m_sharded.log
for n in 2 4 8; do
  P=$((P+1)); timeout 900 $TR --nproc-per-node $n --master-port $P bench.py --gpus $n --config 608x320 --no-cpu-baseline > gpurun_out/r2m_608x320_${n}gpu.json 2> gpurun_out/r2m_608x320_${n}gpu.err
  cut -c1-200 gpurun_out/r2m_608x320_${n}gpu.json; echo
done
for n in 8 4 2; do
  P=$((P+1)); timeout 900 $TR --nproc-per-node $n --master-port $P bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_headline_${n}gpu.json 2> gpurun_out/r2m_headline_${n}gpu.err
  cut -c1-200 gpurun_out/r2m_headline_${n}gpu.json; echo
done
