#!/bin/bash
# auto chain schedule: where does the list schedule pay?  (size, batch) sweep with FVY_CHAIN_SCHED = 0 / 1
run() { # name, args..., env via SCHED
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2q_$1.json")); r = d["roofline"]
    print("$1: value %.0f ms %.3f | fwd %.3f | alone fwd %.3f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["forward_ms_alone"]))
except Exception as e:
    print("$1 failed", e)
PY
}
for rep in 1 2; do
for cfg in "608 40" "608 16" "416 80" "416 160" "416 16" "608 80" "320 40"; do
  set -- $cfg
  for s in 0 1; do
    FVY_CHAIN_SCHED=$s timeout 300 python bench.py --size $1 --batch $2 --steps 20 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2q_$1_$2_s${s}_$rep.json 2>> gpurun_out/r2q_bench.err
    run $1_$2_s${s}_$rep
  done
done
done
tail -3 gpurun_out/r2q_bench.err
