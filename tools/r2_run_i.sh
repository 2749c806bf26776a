#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "wide_1x1" -p no:cacheprovider > gpurun_out/r2i_test.log 2>&1; echo "wide1x1 test rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2i_test.log | head -20
for m in 0 1; do for sc in 0 1; do
  FVY_CHAIN_128=$m FVY_CHAIN_SCHED=$sc timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2i_bench_c${m}_s${sc}.json 2>> gpurun_out/r2i_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2i_bench_c${m}_s${sc}.json")); r = d["roofline"]
    print("chain128 $m sched $sc: value %.0f ms %.3f | fwd %.3f | alone fwd %.3f post %.3f | launches/step %.0f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"], d["gpu_launches_per_step"]))
except Exception as e:
    print("variant $m $sc failed", e)
PY
done; done
FVY_CHAIN_128=1 timeout 600 python tools/gpu_check.py --tile-n 256 > gpurun_out/r2i_check.log 2>&1; grep -E "conv_1[3-9] |conv_2|conv_3[0-5] |forward batch|sum of isolated" gpurun_out/r2i_check.log | head -30
tail -3 gpurun_out/r2i_bench.err
