#!/bin/bash
# round 2, run D: whole GPU suite per file, then A/B of the overlap / chain-schedule knobs on the headline bench
for f in test_gpu_parity test_gpu_round2 test_train; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/r2d_$f.log 2>&1; echo "$f rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2d_$f.log | tail -25
done
for ov in 1 0; do for sc in 0 1; do
  FVY_OVERLAP_POST=$ov FVY_CHAIN_SCHED=$sc python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2d_bench_ov${ov}_sc${sc}.json 2>> gpurun_out/r2d_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2d_bench_ov${ov}_sc${sc}.json")); r = d["roofline"]
print("overlap $ov sched $sc: value %.0f ms %.3f | fwd %.3f post %.3f | alone fwd %.3f post %.3f | e2e %.0f (f32 %.0f)" % (d["value"], d["ms_per_step"], r["forward_ms"], r["postprocess_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"], d["e2e"]["value"], d["e2e"]["float32_frames"]["value"]))
PY
done; done
tail -3 gpurun_out/r2d_bench.err
