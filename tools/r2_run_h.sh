#!/bin/bash
# fused stem: parity, timing A/B, then one ncu --set full capture of the fused kernel (after the plain runs exited 0)
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "fused_stem" -p no:cacheprovider > gpurun_out/r2h_fused.log 2>&1; echo "fused test rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2h_fused.log | head -30
for m in 0 1; do
  FVY_FUSE_STEM=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2h_bench_fuse$m.json 2>> gpurun_out/r2h_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2h_bench_fuse$m.json")); r = d["roofline"]
    print("fuse $m: value %.0f ms %.3f | fwd %.3f post %.3f | alone fwd %.3f post %.3f | launches/step %.0f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["postprocess_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"], d["gpu_launches_per_step"]))
except Exception as e:
    print("fuse $m failed", e)
PY
done
cat > /tmp/fwd3.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine
import torch
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
xd = torch.from_numpy(synth.images(40, 416, 416, 1)).cuda()
for _ in range(3):
    eng.forward(xd, want_outputs=False)
PY
FVY_FUSE_STEM=1 FVY_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:stem_conv1_fused -s 1 -c 1 -o gpurun_out/r2h_fused -f python /tmp/fwd3.py > gpurun_out/r2h_ncu.log 2>&1
ncu -i gpurun_out/r2h_fused.ncu-rep --page raw --csv > gpurun_out/r2h_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/r2h_fused.ncu-rep --page source --csv > gpurun_out/r2h_fused_source.csv 2>/dev/null
rm -f gpurun_out/r2h_fused.ncu-rep
ls -la gpurun_out/r2h_fused_raw.csv gpurun_out/r2h_fused_source.csv
tail -3 gpurun_out/r2h_bench.err
