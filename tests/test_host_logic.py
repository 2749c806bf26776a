"""Host-side logic that needs no GPU: layer tables, weight-stream layout, sharding (incl. a
world_size-2 gloo run), BoundBox semantics, config loading."""
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from face_vijnana_yolov3_b200 import arch, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_length_and_layout():
    specs = arch.yolo3_table(1)
    s = synth.darknet_stream(specs, 3, synth.INIT_BN_EXERCISING)
    assert s.dtype == np.float32 and s.size == arch.n_params(specs) == 61576342
    c0 = specs[0]
    beta, gamma = s[:32], s[32:64]                      # yolov3_detect.py:97-101: beta, gamma, mean, var, then kernel
    assert np.all((gamma > 0.74) & (gamma < 1.26)) and np.abs(beta).max() < 1.0
    assert c0.n_params == 4 * 32 + 32 * 3 * 9
    k = synth.darknet_stream(specs[:1], 3, synth.INIT_KERAS_DEFAULT)
    assert np.all(k[:32] == 0) and np.all(k[32:64] == 1) and np.all(k[96:128] == 1)


def test_weight_file_header_roundtrip(tmp_path):
    import importlib
    specs = arch.yolo3_table(1)[:3]
    s = synth.darknet_stream(specs, 0)
    p = tmp_path / "w.weights"
    p.write_bytes(synth.darknet_file_bytes(s, 0, 2, 0, 123))
    raw = p.read_bytes()
    assert struct.unpack("iii", raw[:12]) == (0, 2, 0) and len(raw) == 12 + 8 + 4 * s.size   # 8-byte `seen` for version >= 0.2
    p.write_bytes(synth.darknet_file_bytes(s, 0, 1, 0, 5))
    assert len(p.read_bytes()) == 12 + 4 + 4 * s.size
    yd = importlib.import_module("face_vijnana_yolov3_b200.space.yolov3_detect")
    wr = yd.WeightReader(str(p))
    assert np.array_equal(wr.read_bytes(s.size), s)


def test_shard_bounds_cover_batch_contiguously():
    for B in (0, 1, 7, 40, 320):
        for N in (1, 2, 4, 8):
            b = shard.shard_bounds(B, N)
            assert b[0][0] == 0 and b[-1][1] == B
            assert all(b[i][1] == b[i + 1][0] for i in range(N - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in shard.shard_bounds(320, 8)] == [40] * 8


def test_gloo_two_ranks_shard_and_gather(tmp_path):
    """N>1 host path on CPU: two gloo ranks each take their slice; rank 0 gathers results in image order."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, json
sys.path.insert(0, {ROOT!r})
import torch, torch.distributed as dist
from face_vijnana_yolov3_b200 import shard
dist.init_process_group('gloo')
r, n = dist.get_rank(), dist.get_world_size()
lo, hi = shard.shard_bounds(11, n)[r]
mine = [i * i for i in range(lo, hi)]            # stand-in for per-image detections
out = [None] * n
dist.all_gather_object(out, mine)
t = torch.tensor([float(r + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max-over-ranks timing
if r == 0:
    print(json.dumps({{'res': shard.gather_in_image_order(out), 'max': t.item()}}))
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["res"] == [i * i for i in range(11)] and d["max"] == 2.0


def test_boundbox_semantics():
    from face_vijnana_yolov3_b200.space.yolov3_detect import BoundBox
    b = BoundBox(1, 2, 3, 4, objness=0.9, classes=np.array([0.2, 1.5, 0.1], np.float32))
    assert b.get_label() == 1 and b.get_score() == 1.0          # clipped to 1.0 (yolov3_detect.py:155)
    b.classes[1] = 0.0
    assert b.get_score() == 1.0                                 # cached on first call (:151-153)
    assert BoundBox(0, 0, 50, 25).get_relative_bb(100, 100) == (0, 0, 50, 25)


def test_reference_config_schema_loads():
    conf = {"fd_conf": {"mode": "test", "raw_data_path": "x", "test_path": "y", "output_file_path": "o.csv", "multi_gpu": True, "num_gpus": 4,
                        "yolov3_base_model_load": True, "resource_type": "vggface2",
                        "hps": {"lr": 1e-4, "beta_1": 0.99, "beta_2": 0.99, "decay": 0.0, "epochs": 6, "step": 1, "batch_size": 40,
                                "face_conf_th": 0.5, "nms_iou_th": 0.5, "num_cands": 60, "face_region_ratio_th": 0.8},
                        "nn_arch": {"image_size": 416, "bb_info_c_size": 6}, "model_loading": False}}
    from face_vijnana_yolov3_b200.space.face_detection import FaceDetector
    fd = FaceDetector(json.loads(json.dumps(conf))["fd_conf"])   # unknown keys ignored; no GPU touched until detect()
    assert fd.cell_image_size == 32 and FaceDetector.CELL_SIZE == 13 and FaceDetector.MODEL_PATH == "face_detector.h5"
    assert fd._stream.size == 40675942
    with pytest.raises(FileNotFoundError):           # train() reads raw_data_path/training.csv like the reference (:82)
        fd.train(device="cpu")


def test_facedetector_letterbox_geometry_follows_reference():
    """FaceDetector._letterbox_geom (what the file loop hands to fvy_letterbox_u8) = the reference's arithmetic
    (face_detection.py:664-688), restated in oracle/letterbox.py and pinned there on the reference's own loop."""
    from face_vijnana_yolov3_b200.space.face_detection import FaceDetector
    from oracle import letterbox as LB
    fd = FaceDetector.__new__(FaceDetector)
    rng = np.random.default_rng(2)
    for size in (416, 608):
        fd.nn_arch = {"image_size": size, "bb_info_c_size": 6}
        for _ in range(200):
            w, h = int(rng.integers(16, 3000)), int(rng.integers(16, 3000))
            w_p, h_p, pad_t, pad_b, pad_l, pad_r = LB.geometry(w, h, size)
            assert fd._letterbox_geom(w, h) == (w_p, h_p, pad_t, pad_l)
            assert pad_t + h_p + pad_b == size and pad_l + w_p + pad_r == size


def test_traffic_profile_is_stamped_with_a_source_hash():
    """profiles/conv_traffic.json (what bench.py's roofline.traffic reads) names the library sources it was captured on; bench.py
    recomputes that hash from csrc/ + include/fvy.h.  A mismatch only means the evidence is older than the sources (bench.py then
    reports traffic_is_this_build = false), so it is a skip, not a failure."""
    import importlib.util
    import json
    import os
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    src = bench._src_sha16()
    assert isinstance(src, str) and len(src) == 16 and int(src, 16) >= 0
    assert src == bench._src_sha16()                      # deterministic
    tr = json.load(open(os.path.join(root, "profiles", "conv_traffic.json")))
    assert tr["dram_bytes_per_step"] > 0 and tr["launches"] == 44 and (tr["batch"], tr["net"]) == (40, 416)
    if tr.get("libfvy_src_sha16") != src:
        pytest.skip("profiles/conv_traffic.json was captured on older sources than this tree")
