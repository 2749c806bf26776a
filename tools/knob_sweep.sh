#!/bin/bash
# Forward time (batch 40 @416, 30 forwards after 10 warm-ups) under different experiment knobs; two passes to see the drift.
cat > /tmp/fwdt.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
xd = torch.from_numpy(synth.images(40, 416, 416, 1)).cuda()
for _ in range(10): eng.forward(xd, want_outputs=False)
ts = []
for _ in range(30):
    eng.forward(xd, want_outputs=False); ts.append(eng.last_timing()[0])
print("forward ms median %.3f min %.3f" % (float(np.median(ts)), min(ts)))
PY
for rep in 1 2; do
for cfg in "FVY_X=0" "FVY_NB_RES=3" "FVY_NB=2" "FVY_FLAGS_MAX_TILES=16" "FVY_FLAGS_MAX_TILES=64" "FVY_CHAIN_A=5 FVY_CHAIN_B=5" "FVY_CHAIN_NB=4 FVY_CHAIN_B=5" "FVY_B3_MAX=49152"; do
  echo "== $cfg: $(env $cfg python /tmp/fwdt.py 2>&1 | tail -1)"
done; done
