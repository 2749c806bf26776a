#!/bin/bash
# compact geometry: new bit-identity test in modes 1 and 2, A/B of mode 2, role counters of the 1x1 layers
for m in 1 2; do
FVY_COMPACT=$m timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "shared_halo" -p no:cacheprovider > gpurun_out/r2k_test_$m.log 2>&1; echo "compact=$m test rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2k_test_$m.log | head -10
done
FVY_COMPACT=2 timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2k_pytest.log 2>&1; echo "pytest (mode 2) rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2k_pytest.log | head -20
for rep in 1 2; do for m in 1 2; do
  FVY_COMPACT=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2k_bench_c${m}_${rep}.json 2>> gpurun_out/r2k_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2k_bench_c${m}_${rep}.json")); r = d["roofline"]
    print("compact $m: value %.0f ms %.3f | fwd %.3f | alone fwd %.3f post %.3f | e2e %.0f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"], d["e2e"]["value"]))
except Exception as e:
    print("variant $m failed", e)
PY
done; done
FVY_DBG=1 timeout 300 python tools/run_layer.py --layers 10,5,2,27,28,46,49 --iters 5 2>&1 | grep "fvy dbg" > gpurun_out/r2k_dbg.log
cat gpurun_out/r2k_dbg.log
tail -3 gpurun_out/r2k_bench.err
