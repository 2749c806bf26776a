"""Pins the oracle (oracle/postproc_c.c) against the reference: (1) the golden vectors that
tools/make_golden.py produced by EXECUTING the reference's own functions, and (2) when
/root/reference is present (build container), the live reference functions themselves.

The reference evaluates sigmoid/exp with NumPy's float32 SIMD exp (<= a few ulp, platform
dependent); the oracle uses the correctly rounded float32 value.  Scores therefore agree to a few
ulp, while candidate sets, integer boxes and the NMS outcome agree exactly on these vectors.
"""
import json
import os

import numpy as np
import pytest

from oracle import postproc as P, ref_loader as R


def _ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


@pytest.mark.parametrize("tag", ["a", "b", "c", "mc"])      # mc: nb_class = 3, pins the per-class loop of do_nms (:431-444)
def test_yolo3_post_against_reference_vectors(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"post_yolo3_{tag}.npz"))
    outs = [g["out0"], g["out1"], g["out2"]]
    d = P.decode_image(outs, obj_thresh=float(g["obj_thresh"]), arith=P.ARITH_F32)   # fixtures were made under NumPy >= 2
    assert len(d["cell"]) == g["nbox"].shape[0]                                       # same candidate set, same order
    assert _ulp_diff(d["objness"], g["objness"]).max() <= 4
    assert _ulp_diff(d["classes"], g["classes_before"]).max() <= 4
    assert np.abs(d["box"] - g["nbox"]).max() <= 1e-6
    ib = P.correct_yolo_boxes(d["box"], int(g["image_hw"][0]), int(g["image_hw"][1]), 416, 416, P.ARITH_F32)
    mism = (ib != g["ibox"]).mean()
    assert mism <= 1e-3, f"{mism:.2e} of integer coordinates differ"                    # exp ulp straddling an integer
    # NMS: bit-exact when fed the reference's own candidates (boxes AND scores)
    cls = P.do_nms(g["ibox"], g["classes_before"], float(g["nms_thresh"]))
    assert np.array_equal(cls, g["classes_after"])


def test_fd6_detect_against_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "post_fd6.npz"))
    hps = json.loads(str(g["hps"]))
    for k in range(4):
        ib, sc, cell = P.fd6_detect(g["maps"][k], 416, hps[k]["face_conf_th"], hps[k]["nms_iou_th"], hps[k]["num_cands"], P.ARITH_F32)
        assert np.array_equal(ib, g[f"ibox{k}"]), k
        assert _ulp_diff(sc, g[f"score{k}"]).max() <= 4


def test_iou_known_answers(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "iou_cases.json")))
    for c in cases:
        v = P.bbox_iou(c["a"], c["b"])
        if c["iou"] is None:
            assert np.isnan(v)
        else:
            assert v == c["iou"], c                                                      # IEEE double, bit-exact


def test_arith_modes_differ_only_at_integer_crossings():
    rng = np.random.default_rng(0)
    outs = [rng.standard_normal((g, g, 18)).astype(np.float32) for g in (13, 26, 52)]
    a = P.decode_image(outs, arith=P.ARITH_F64)
    b = P.decode_image(outs, arith=P.ARITH_F32)
    assert np.array_equal(a["cell"], b["cell"])
    ia = P.correct_yolo_boxes(a["box"], 360, 640, 416, 416, P.ARITH_F64)
    ib = P.correct_yolo_boxes(b["box"], 360, 640, 416, 416, P.ARITH_F32)
    assert (ia != ib).mean() < 1e-3 and np.abs(ia - ib).max() <= 1


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_live_reference_matches_oracle():
    from face_vijnana_yolov3_b200 import synth
    Y = R.load_yolov3_detect()
    outs = synth.head_logits(1, 416, 416, 1, seed=77)
    boxes = []
    for i in range(3):
        boxes += Y.decode_netout(outs[i][0].copy(), list(P.REF_ANCHORS[i]), i, 0.5, 416, 416)
    d = P.decode_image([o[0] for o in outs], arith=P.ARITH_F32)
    assert len(boxes) == len(d["cell"])
    Y.correct_yolo_boxes(boxes, 300, 500, 416, 416)
    rib = np.array([[b.xmin, b.ymin, b.xmax, b.ymax] for b in boxes], np.int64)
    ib = P.correct_yolo_boxes(d["box"], 300, 500, 416, 416, P.ARITH_F32)
    assert (rib != ib).mean() <= 1e-3
    rc = np.array([b.classes[0] for b in boxes], np.float32)
    sub = slice(0, 600)                                                                   # the reference NMS is O(n^2) Python
    bsub = boxes[sub]
    Y.do_nms(bsub, 0.45)
    cls = P.do_nms(rib[sub], rc[sub], 0.45)
    assert np.array_equal(np.array([b.classes[0] for b in bsub], np.float32), cls[:, 0])


def _write_map_case(g, tag, d):
    gt, sol = os.path.join(d, "gt.csv"), os.path.join(d, "sol.csv")
    open(gt, "w").write(str(g[f"gt_csv_{tag}"])); open(sol, "w").write(str(g[f"sol_csv_{tag}"]))
    return gt, sol


@pytest.mark.parametrize("tag", ["f", "i"])
@pytest.mark.filterwarnings("ignore")
def test_map_fd_oracle_against_reference_vectors(golden_dir, tmp_path, tag):
    """oracle/map_fd.py vs the outputs of the reference's own cal_mAP_fd (evaluate.py:27-127): ps, rs and mAP bit for bit."""
    from oracle import map_fd as M
    g = np.load(os.path.join(golden_dir, "map_fd.npz"))
    gt, sol = _write_map_case(g, tag, str(tmp_path))
    for th in (0.5, 0.75):
        ps, rs, mAP = M.cal_mAP_fd(gt, sol, th)
        k = int(th * 100)
        assert np.array_equal(ps, g[f"ps_{tag}_{k}"]) and np.array_equal(rs, g[f"rs_{tag}_{k}"]) and mAP == float(g[f"mAP_{tag}_{k}"])


@pytest.mark.skipif(not R.available(), reason="needs /root/reference (build container only)")
@pytest.mark.filterwarnings("ignore")
def test_map_fd_oracle_against_live_reference(golden_dir, tmp_path):
    from oracle import map_fd as M
    E = R.load_evaluate()
    g = np.load(os.path.join(golden_dir, "map_fd.npz"))
    gt, sol = _write_map_case(g, "f", str(tmp_path))
    for th in (0.5, 0.6, 0.9):
        a = E.cal_mAP_fd(gt, sol, th); b = M.cal_mAP_fd(gt, sol, th)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
