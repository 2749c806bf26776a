cat > /tmp/fwdt.py <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
xd = torch.from_numpy(synth.images(40, 416, 416, 1)).cuda()
for _ in range(4): eng.forward(xd, want_outputs=False)
PY
for f in 1 0; do echo "== FVY_FLAGS=$f"; FVY_TRACE=1 FVY_GRAPH=0 FVY_FLAGS=$f python /tmp/fwdt.py 2>&1 | grep "trace" | tail -75 > gpurun_out/trace$f.log; tail -45 gpurun_out/trace$f.log | head -30; done
