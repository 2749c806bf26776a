#!/bin/bash
# Layer chains with the static rotation vs the host list schedule (FVY_CHAIN_SCHED=1): bit-equality against the un-chained path, forward time.
T="tests/test_gpu_parity.py::test_full_size_batch_invariance_and_tile_dependency_equivalence"
echo "== default"; timeout 120 python -m pytest $T -x -q 2>&1 | tail -2
echo "== FVY_CHAIN_SCHED=1"; FVY_CHAIN_SCHED=1 timeout 120 python -m pytest $T -x -q 2>&1 | tail -4
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
for cfg in "FVY_CHAIN_SCHED=0" "FVY_CHAIN_SCHED=1" "FVY_CHAIN_SCHED=0" "FVY_CHAIN_SCHED=1"; do echo "== $cfg: $(env $cfg timeout 60 python /tmp/fwdt.py 2>&1 | tail -1)"; done
