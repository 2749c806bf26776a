"""The forward-pass restatement (oracle/darknet_ref.py) against the only fixtures the reference
ships for it: the Keras summary() dump of the Darknet-53 base (analysis/face_recog_analysis.ipynb)."""
import json
import os

import numpy as np

from face_vijnana_yolov3_b200 import arch, synth
from oracle import darknet_ref as D


def test_layer_shapes_and_params_match_keras_summary(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "darknet53_summary.json")))
    convs = {r[0]: r for r in g["rows"] if r[1] == "Conv2D"}
    bns = {r[0]: r for r in g["rows"] if r[1] == "BatchNormalization"}
    adds = [r for r in g["rows"] if r[1] == "Add"]
    table = [c for c in arch.fd6_table() if c.idx <= 73]
    assert len(table) == len(convs) == 52 and len(adds) == 23
    for c in table:
        name, _, h, w, ch, npar = convs[f"conv_{c.idx}"]
        assert (h, w, ch) == (416 >> c.level, 416 >> c.level, c.cout)
        assert npar == c.n_kernel                                   # use_bias=False with bnorm (yolov3_detect.py:211)
        assert bns[f"bnorm_{c.idx}"][5] == 4 * c.cout
    assert sum(1 for c in table if c.res is not None) == 23
    assert arch.n_params(table) == g["total"] == 40620640
    assert sum(4 * c.cout for c in table) // 2 == g["non_trainable"]   # moving mean + variance
    # the oracle's own literal table agrees with the product table (independent restatements)
    oc = D.conv_list(18, fd6=True)
    assert [(i, cin, cv["filter"], cv["kernel"], cv["stride"]) for i, cin, cv in oc] == \
           [(c.idx, c.cin, c.cout, c.k, c.stride) for c in arch.fd6_table()]
    oc = D.conv_list(18)
    assert [(i, cin, cv["filter"], cv["kernel"], cv["stride"], cv["bnorm"]) for i, cin, cv in oc] == \
           [(c.idx, c.cin, c.cout, c.k, c.stride, c.bn) for c in arch.yolo3_table(1)]


def test_param_counts_and_macs():
    assert arch.n_params(arch.yolo3_table(1)) == 61576342
    assert arch.n_params(arch.yolo3_table(80)) == 62001757
    assert arch.macs(arch.yolo3_table(1), 416, 416) == 32644937728
    assert arch.macs(arch.yolo3_table(1), 608, 608) == 69732677632
    assert arch.n_params(arch.fd6_table()) == 40620640 + 55302


def test_forward_shapes_small_input():
    specs = arch.yolo3_table(1)
    stream = synth.darknet_stream(specs, 0, synth.INIT_KERAS_DEFAULT)
    x = synth.images(1, 64, 96, 0)
    outs = D.forward(stream, x, 1)
    assert [o.shape for o in outs] == [(1, 2, 3, 18), (1, 4, 6, 18), (1, 8, 12, 18)]
    assert all(np.isfinite(o).all() for o in outs)
    emu = D.forward(stream, x, 1, emulate_bf16=True)
    for a, b in zip(emu, outs):
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 2e-2
    fd = D.forward(synth.darknet_stream(arch.fd6_table(), 0), synth.images(1, 416, 416, 0)[:, :64, :64], fd6=True)
    assert fd.shape == (1, 2, 2, 6)
