cat > /tmp/fwd1.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine
import torch
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
xd = torch.from_numpy(synth.images(40, 416, 416, 1)).cuda()
for _ in range(2):
    eng.forward(xd, want_outputs=False)
PY
FVY_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:stem_ -s 1 -c 1 -o gpurun_out/prof_stem -f python /tmp/fwd1.py > gpurun_out/prof_stem.log 2>&1
ncu -i gpurun_out/prof_stem.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Issue Slots|Issued Ipc|Eligible|Stall|stall|Registers|Occupancy|Local|Bank|Warp Cycles Per Issued|L1/TEX Hit" | head -40
ncu -i gpurun_out/prof_stem.ncu-rep --page source --csv --print-source sass 2>/dev/null > gpurun_out/prof_stem_src.csv
