#!/bin/bash
# per-kernel times of the post-processing kernels of one detect call (ncu launch list; cold-cache, serialised)
cat > /tmp/fwd2.py <<'PY'
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
    d, c = eng.detect(xd, pp=post_params(0.5, 0.45), image_hw=hw, max_out=4225)
print("kept per image", c.mean(), "launches", eng.launch_count)
import time
for tag in ("a", "b"):
    t = []
    for _ in range(20):
        eng.detect(xd, pp=post_params(0.5, 0.45), image_hw=hw, max_out=4225)
        t.append(eng.last_timing())
    print("alone fwd/post ms (median of 20):", np.median([a for a, b in t]), np.median([b for a, b in t]))
PY
python /tmp/fwd2.py
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_yolo|sort_scores|nms_mask|nms_sweep|assemble" -c 15 --csv --log-file gpurun_out/r2e_post_launches.csv python /tmp/fwd2.py > gpurun_out/r2e_ncu.log 2>&1
grep -v "^==" gpurun_out/r2e_post_launches.csv | cut -d, -f5,9,10,15 | tail -16
FVY_SWEEP_BIG=1 python /tmp/fwd2.py
