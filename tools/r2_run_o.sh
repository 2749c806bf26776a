#!/bin/bash
# knob A/B on the compact-geometry build (two repetitions each, interleaved)
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2o_$name.json 2>> gpurun_out/r2o_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2o_$name.json")); r = d["roofline"]
    print("$name: value %.0f ms %.3f | fwd %.3f | alone fwd %.3f post %.3f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"]))
except Exception as e:
    print("$name failed", e)
PY
}
for rep in 1 2; do
  run base_$rep FVY_X=0
  run sched_$rep FVY_CHAIN_SCHED=1
  run pair1x1_$rep FVY_CTA2_128_1X1=1
  run chain128_$rep FVY_CHAIN_128=1
  run chain128s_$rep FVY_CHAIN_128=1 FVY_CHAIN_SCHED=1
  run nooverlap_$rep FVY_OVERLAP_POST=0
done
tail -3 gpurun_out/r2o_bench.err
