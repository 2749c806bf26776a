#!/bin/bash
# Quick A/B on the GPU box: isolated layer times, forward time, GPU parity suite.  usage: quick_ab.sh [layers]
LAYERS=${1:-1,2,3,4,5,6,10,11,27,28}
python tools/run_layer.py --layers $LAYERS --iters 10 2>&1 | grep -v "^fvy" | sed -e "s/{'idx': \([0-9]*\),.*'stages': \([0-9]*\).*}/conv_\1 st=\2/"
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
python /tmp/fwdt.py 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
