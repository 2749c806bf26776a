#!/bin/bash
# DRAM traffic of every conv launch of one forward (batch 40 @416): metrics-only ncu pass -> gpurun_out/traffic.csv
cat > /tmp/fwd2.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
x = synth.images(40, 416, 416, 1)
import torch
xd = torch.from_numpy(x).cuda()
for _ in range(2):
    eng.forward(xd, want_outputs=False)
PY
FVY_GRAPH=0 python /tmp/fwd2.py || exit 1
FVY_GRAPH=0 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv_igemm|stem_" -s 75 -c 75 --csv --log-file gpurun_out/traffic.csv python /tmp/fwd2.py > gpurun_out/traffic_ncu.log 2>&1
tail -3 gpurun_out/traffic.csv
