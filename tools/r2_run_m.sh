#!/bin/bash
# mask kernel v3 (packed-half pre-filter, per-32 flags from the sort kernel): parity suites, headline + stress benches, per-kernel times
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2m2_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2m2_pytest.log | head -20
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r2m2_bench.json 2> gpurun_out/r2m2_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2m2_bench.json")); r = d["roofline"]
print("headline: value %.0f ms %.3f | fwd %.3f | alone fwd %.3f post %.3f | e2e %.0f" % (d["value"], d["ms_per_step"], r["forward_ms"], r["forward_ms_alone"], r["postprocess_ms_alone"], d["e2e"]["value"]))
PY
timeout 300 python bench.py --config stress --no-cpu-baseline > gpurun_out/r2m2_stress.json 2>> gpurun_out/r2m2_bench.err; cut -c1-200 gpurun_out/r2m2_stress.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2m2_stress_launches.csv \
    python bench.py --config stress --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2m2_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2m2_stress_launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[-5:]:
    print("stress", r[4][:40], r[-1])
PY
cat > /tmp/det.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine, post_params
import torch, numpy as np
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
xd = torch.from_numpy(synth.images(40, 416, 416, 1)).cuda()
hw = np.array([[416, 416]] * 40, np.int32)
for _ in range(3):
    eng.detect(xd, pp=post_params(0.5, 0.45), image_hw=hw, max_out=4225)
PY
python /tmp/det.py && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_yolo|sort_scores|nms_mask|nms_sweep|assemble" -c 15 --csv --log-file gpurun_out/r2m2_post_launches.csv python /tmp/det.py > gpurun_out/r2m2_ncu2.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2m2_post_launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[-5:]:
    print("headline", r[4][:40], r[-1])
PY
tail -3 gpurun_out/r2m2_bench.err
