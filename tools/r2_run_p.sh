#!/bin/bash
# chain list schedule as the default: whole GPU suite with it, then a sweep of the schedule's cost model
FVY_CHAIN_SCHED=1 timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2p_pytest.log 2>&1; echo "pytest (sched) rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2p_pytest.log | head -20
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2p_$name.json 2>> gpurun_out/r2p_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2p_$name.json")); r = d["roofline"]
    print("$name: value %.0f ms %.3f | fwd %.3f | alone fwd %.3f post %.3f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"]))
except Exception as e:
    print("$name failed", e)
PY
}
for rep in 1 2; do
  run s_5_10_$rep FVY_CHAIN_SCHED=1
  run s_2_4_$rep FVY_CHAIN_SCHED=1 FVY_SCHED_LOAD=2 FVY_SCHED_DRAIN=4
  run s_3_6_$rep FVY_CHAIN_SCHED=1 FVY_SCHED_LOAD=3 FVY_SCHED_DRAIN=6
  run s_8_16_$rep FVY_CHAIN_SCHED=1 FVY_SCHED_LOAD=8 FVY_SCHED_DRAIN=16
  run s_5_20_$rep FVY_CHAIN_SCHED=1 FVY_SCHED_LOAD=5 FVY_SCHED_DRAIN=20
  run s_0_0_$rep FVY_CHAIN_SCHED=1 FVY_SCHED_LOAD=0 FVY_SCHED_DRAIN=0
  run s_maxlen8_$rep FVY_CHAIN_SCHED=1 FVY_CHAIN_MAXLEN=8
done
for c in 608x320; do
  for s in 0 1; do
    FVY_CHAIN_SCHED=$s timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2p_${c}_s$s.json 2>> gpurun_out/r2p_bench.err
    python - <<PY
import json
d = json.load(open("gpurun_out/r2p_${c}_s$s.json")); print("$c sched $s: value %.0f ms %.3f" % (d["value"], d["ms_per_step"]))
PY
  done
done
tail -3 gpurun_out/r2p_bench.err
