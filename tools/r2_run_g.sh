#!/bin/bash
# fused stem: parity first (bounded), then timing A/B
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "fused_stem" -p no:cacheprovider > gpurun_out/r2g_fused.log 2>&1; echo "fused test rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2g_fused.log | head -30
for m in 0 1; do
  FVY_FUSE_STEM=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2g_bench_fuse$m.json 2>> gpurun_out/r2g_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2g_bench_fuse$m.json")); r = d["roofline"]
    print("fuse $m: value %.0f ms %.3f | fwd %.3f post %.3f | alone fwd %.3f post %.3f | launches/step %.0f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["postprocess_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"], d["gpu_launches_per_step"]))
except Exception as e:
    print("fuse $m failed", e)
PY
done
tail -5 gpurun_out/r2g_bench.err
