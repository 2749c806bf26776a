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
for rep in 1 2; do for cfg in "FVY_CHAIN=0" "FVY_CHAIN=1" "FVY_CHAIN=1 FVY_CHAIN_A=3 FVY_CHAIN_B=7" "FVY_CHAIN=1 FVY_CHAIN_MAXLEN=4"; do echo "== $cfg"; env $cfg python /tmp/fwdt.py 2>&1 | tail -1; done; done
