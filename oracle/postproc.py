"""ORACLE — test infrastructure only.  ctypes wrapper around ``postproc_c.c``.

Restates (array-in / array-out) the reference's ``decode_netout``, ``correct_yolo_boxes``,
``bbox_iou``, ``do_nms(_v2)`` (``src/space/yolov3_detect.py:165-194,335-458``) and the decode /
select halves of ``FaceDetector.detect`` (``src/space/face_detection.py:899-947``).

Parity pin: ``tests/test_oracle_pin.py`` runs the REAL reference functions (``ref_loader``) on
seeded inputs in the build container and compares; ``tests/golden/*.npz`` hold their outputs
for the GPU box where /root/reference does not exist (generator: ``tools/make_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_postproc.so")

ARITH_F64 = 0   # NumPy 1.x promotion (reference's pinned environment)
ARITH_F32 = 1   # NumPy >= 2 promotion (reference executed in this container)

REF_ANCHORS = ((116, 90, 156, 198, 373, 326), (30, 61, 62, 45, 59, 119), (10, 13, 16, 30, 33, 23))  # yolov3_detect.py:558-560
REF_ANCHOR_MASK = (0b010, 0b101, 0b010)   # yolov3_detect.py:354-362 (fork-specific)
ALL_ANCHOR_MASK = (0b111, 0b111, 0b111)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "postproc_c.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_postproc.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        f32p, f64p, i32p, i64p = (C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int64))
        L.orc_sigmoid.restype = C.c_float; L.orc_sigmoid.argtypes = [C.c_float]
        L.orc_bbox_iou_i.restype = C.c_double; L.orc_bbox_iou_i.argtypes = [i64p, i64p]
        L.orc_decode_netout.restype = C.c_int
        L.orc_decode_netout.argtypes = [f32p, C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.c_double, C.c_int, C.c_int,
                                        C.c_int, f64p, f32p, f32p, i32p, C.c_int]
        L.orc_correct_yolo_boxes.restype = None
        L.orc_correct_yolo_boxes.argtypes = [f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64p]
        L.orc_do_nms.restype = None
        L.orc_do_nms.argtypes = [i64p, C.c_int, C.c_int, C.c_double, f32p]
        L.orc_fd6_decode.restype = C.c_int
        L.orc_fd6_decode.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, i64p, f32p, f32p, i32p]
        L.orc_fd6_select.restype = C.c_int
        L.orc_fd6_select.argtypes = [f32p, C.c_int, C.c_int, i32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def sigmoid(x):
    L = lib()
    x = np.asarray(x, np.float32)
    return np.array([L.orc_sigmoid(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)


def bbox_iou(a, b) -> float:
    a = np.ascontiguousarray(a, np.int64); b = np.ascontiguousarray(b, np.int64)
    return float(lib().orc_bbox_iou_i(_p(a, C.c_int64), _p(b, C.c_int64)))


def decode_netout(netout, anchors6, anchor_mask3, obj_thresh, net_h, net_w, arith=ARITH_F64):
    """One scale of one image -> dict(box[n,4] f64, objness[n], classes[n,nc], cell[n])."""
    netout = np.ascontiguousarray(netout, np.float32)
    gh, gw, ch = netout.shape
    nc = ch // 3 - 5
    cap = gh * gw * 3
    box = np.empty((cap, 4), np.float64); obj = np.empty(cap, np.float32)
    cls = np.empty((cap, max(nc, 1)), np.float32); cell = np.empty(cap, np.int32)
    anc = np.ascontiguousarray(anchors6, np.int32)
    n = lib().orc_decode_netout(_p(netout, C.c_float), gh, gw, nc, _p(anc, C.c_int), int(anchor_mask3), float(obj_thresh),
                                int(net_h), int(net_w), int(arith), _p(box, C.c_double), _p(obj, C.c_float),
                                _p(cls, C.c_float), _p(cell, C.c_int), cap)
    assert n >= 0
    return dict(box=box[:n].copy(), objness=obj[:n].copy(), classes=cls[:n, :nc].copy(), cell=cell[:n].copy())


def decode_image(netouts, anchors=REF_ANCHORS, anchor_masks=REF_ANCHOR_MASK, obj_thresh=0.5, net_h=416, net_w=416,
                 arith=ARITH_F64):
    """The reference's per-image loop (yolov3_detect.py:596-598): scales 0,1,2 concatenated."""
    parts = [decode_netout(netouts[i], anchors[i], anchor_masks[i], obj_thresh, net_h, net_w, arith) for i in range(3)]
    scale = np.concatenate([np.full(len(p["cell"]), i, np.int32) for i, p in enumerate(parts)])
    out = {k: np.concatenate([p[k] for p in parts]) for k in ("box", "objness", "classes", "cell")}
    out["scale"] = scale
    return out


def correct_yolo_boxes(box, image_h, image_w, net_h, net_w, arith=ARITH_F64):
    box = np.ascontiguousarray(box, np.float64)
    n = box.shape[0]
    ibox = np.empty((n, 4), np.int64)
    lib().orc_correct_yolo_boxes(_p(box, C.c_double), n, int(image_h), int(image_w), int(net_h), int(net_w), int(arith),
                                 _p(ibox, C.c_int64))
    return ibox


def do_nms(ibox, classes, nms_thresh):
    """Returns a copy of ``classes`` with suppressed scores zeroed (reference mutates in place)."""
    ibox = np.ascontiguousarray(ibox, np.int64)
    classes = np.array(classes, np.float32, copy=True, order="C")
    if classes.ndim == 1:
        classes = classes[:, None]
    n, nc = classes.shape
    lib().orc_do_nms(_p(ibox, C.c_int64), n, nc, float(nms_thresh), _p(classes, C.c_float))
    return classes


def fd6_decode(cands, image_size=416, face_conf_th=0.5, arith=ARITH_F64, grid=None):
    cands = np.ascontiguousarray(cands, np.float32)
    grid = grid or cands.shape[0]
    cap = grid * grid
    ibox = np.empty((cap, 4), np.int64); obj = np.empty(cap, np.float32); sc = np.empty(cap, np.float32)
    cell = np.empty(cap, np.int32)
    n = lib().orc_fd6_decode(_p(cands, C.c_float), grid, int(image_size), int(image_size) // 13, float(face_conf_th), int(arith),
                             _p(ibox, C.c_int64), _p(obj, C.c_float), _p(sc, C.c_float), _p(cell, C.c_int))
    return dict(ibox=ibox[:n].copy(), objness=obj[:n].copy(), score=sc[:n].copy(), cell=cell[:n].copy())


def fd6_detect(cands, image_size=416, face_conf_th=0.5, nms_iou_th=0.5, num_cands=60, arith=ARITH_F64):
    """Whole post-predict part of FaceDetector.detect -> (ibox[k,4], score[k], cell[k]) in return order."""
    d = fd6_decode(cands, image_size, face_conf_th, arith)
    n = len(d["cell"])
    if n == 0:
        return np.zeros((0, 4), np.int64), np.zeros(0, np.float32), np.zeros(0, np.int32)
    sc = do_nms(d["ibox"], d["score"], nms_iou_th)[:, 0]
    out = np.empty(max(n, 1), np.int32)
    k = lib().orc_fd6_select(_p(np.ascontiguousarray(sc), C.c_float), n, int(num_cands), _p(out, C.c_int))
    sel = out[:k]
    return d["ibox"][sel], sc[sel], d["cell"][sel]
