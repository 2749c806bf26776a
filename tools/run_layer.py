"""Run selected conv layers in isolation (for ncu captures).  usage: run_layer.py --layers 45,13 --batch 40 --iters 2 [--tile-n 256]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_vijnana_yolov3_b200 import _lib as L, arch, synth
from face_vijnana_yolov3_b200.engine import Engine
import ctypes as C

ap = argparse.ArgumentParser()
ap.add_argument("--layers", default="45")
ap.add_argument("--batch", type=int, default=40)
ap.add_argument("--size", type=int, default=416)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--tile-n", type=int, default=0)
a = ap.parse_args()
eng = Engine(a.size, a.size, nb_class=1, max_batch=a.batch, tile_n_max=a.tile_n)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
x = synth.images(a.batch, a.size, a.size, 1)
eng.forward(x, want_outputs=False)           # fills every activation buffer with real data
infos = eng.layer_infos()
sel = [int(s) for s in a.layers.split(",")]
for li in sel:                               # each: 1 + iters launches, in this order (ncu: -s 75 -c len*(1+iters))
    ms = eng.run_layer(li, a.batch, a.iters)
    print(li, infos[li], f"{ms*1e3:.1f} us", flush=True)
