#!/bin/bash
# mask kernel v3 with 16 resident blocks per SM: post tests, per-kernel times, full capture at the headline configuration
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu -x -k "post or nms or golden or detect or stress or cand or mc or float" -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|^E  |passed|failed|timed out|fvy:" gpurun_out/r2n_pytest.log | head -20
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
python /tmp/det.py && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_yolo|sort_scores|nms_mask|nms_sweep|assemble" -c 15 --csv --log-file gpurun_out/r2n_post_launches.csv python /tmp/det.py > gpurun_out/r2n_ncu2.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2n_post_launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[-5:]:
    print("headline", r[4][:40], r[-1])
PY
ncu --set full --clock-control none --import-source on -k regex:"nms_mask" -s 2 -c 1 -o gpurun_out/r2n_mask -f python /tmp/det.py > gpurun_out/r2n_ncu3.log 2>&1
ncu -i gpurun_out/r2n_mask.ncu-rep --page raw --csv > gpurun_out/r2n_mask_raw.csv 2>/dev/null
ncu -i gpurun_out/r2n_mask.ncu-rep --page source --csv > gpurun_out/r2n_mask_source.csv 2>/dev/null; rm -f gpurun_out/r2n_mask.ncu-rep
timeout 300 python bench.py --config stress --no-cpu-baseline > gpurun_out/r2n_stress.json 2>> gpurun_out/r2n_bench.err; cut -c1-200 gpurun_out/r2n_stress.json
