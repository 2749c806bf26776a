cat > /tmp/tr.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
xd = torch.from_numpy(synth.images(40, 416, 416, 1)).cuda()
for _ in range(3): eng.forward(xd, want_outputs=False)
PY
FVY_GRAPH=0 FVY_TRACE=1 FVY_CHAIN=0 python /tmp/tr.py 2>&1 | grep "^trace" | tail -74 > gpurun_out/trace_nochain.log
FVY_GRAPH=0 FVY_TRACE=1 python /tmp/tr.py 2>&1 | grep "^trace" | tail -40 > gpurun_out/trace_chain.log
wc -l gpurun_out/trace_*.log
