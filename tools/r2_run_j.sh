#!/bin/bash
# compact (shared-halo) geometry of the narrow levels: parity suites, then A/B against the legacy geometry
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2j_pytest.log | head -20
for rep in 1 2; do for m in 0 1; do
  FVY_COMPACT=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2j_bench_c${m}_${rep}.json 2>> gpurun_out/r2j_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2j_bench_c${m}_${rep}.json")); r = d["roofline"]
    print("compact $m: value %.0f ms %.3f | fwd %.3f | alone fwd %.3f post %.3f | e2e %.0f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"], d["e2e"]["value"]))
except Exception as e:
    print("variant $m failed", e)
PY
done; done
timeout 600 python tools/gpu_check.py --tile-n 256 > gpurun_out/r2j_check.log 2>&1; grep -E "sum of isolated|heads_vs" -A3 gpurun_out/r2j_check.log | head -30
tail -3 gpurun_out/r2j_bench.err
