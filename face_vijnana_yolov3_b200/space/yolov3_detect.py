"""Drop-in for the reference's ``src/space/yolov3_detect.py`` hot-path surface, backed by libfvy.so.

Same names, argument meaning and in-place behaviour as the reference:

    make_yolov3_model()  -> object with .predict(x)            reference :217-311, :593
    WeightReader(path).load_weights(model)                     reference :67-121
    decode_netout(netout, anchors, anchor_idx, obj_thresh, net_h, net_w) -> [BoundBox]   :335-387
    correct_yolo_boxes(boxes, image_h, image_w, net_h, net_w)  (in place)                 :389-404
    correct_yolo_boxes_v2(boxes, image_size, net_h, net_w)     (in place)                 :406-424
    do_nms(boxes, nms_thresh) / do_nms_v2(boxes, nms_thresh)   (in place zeroing)         :426-458
    bbox_iou(box1, box2) -> float                                                        :183-194
    BoundBox(xmin, ymin, xmax, ymax, objness, classes, anchor, subject_id)                :126-163

All arithmetic runs on the GPU through the C ABI; this file only marshals Python objects.  There is
no CPU fallback: without libfvy.so / a B200 the calls raise.  Out of scope (host image I/O, drawing,
COCO demo): preprocess_input, draw_boxes*, get_person_boxes, _main_.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

from .. import _lib as L
from .. import arch, synth
from ..engine import Engine, REF_ANCHORS, post_params

# Arithmetic of the scalar decode math: ARITH_F64 reproduces the reference's pinned NumPy 1.x
# environment (python-int (+) np.float32 -> float64); ARITH_F32 reproduces NumPy >= 2.
DECODE_ARITH = L.ARITH_F64


class BoundBox:
    """Box record of the reference (:126-163); same attributes and lazy label/score."""

    def __init__(self, xmin, ymin, xmax, ymax, objness=None, classes=None, anchor=None, subject_id=-1):
        self.xmin = xmin
        self.ymin = ymin
        self.xmax = xmax
        self.ymax = ymax
        self.objness = objness
        self.classes = classes
        self.anchor = anchor
        self.subject_id = subject_id
        self.label = -1
        self.score = -1

    def get_label(self):
        if self.label == -1:
            self.label = np.argmax(self.classes)
        return self.label

    def get_score(self):
        if self.score == -1:
            self.score = self.classes[self.get_label()]
        return np.min([self.score, 1.0])

    def get_relative_bb(self, width, height):
        left = int(self.xmin / width * 100.)
        top = int(self.ymin / height * 100.)
        width = int((self.xmax - self.xmin) / width * 100.)
        height = int((self.ymax - self.ymin) / height * 100.)
        return (left, top, width, height)


# ----------------------------------------------------------------------------------------------
# shared post-processing handles (FVY_HEAD_NONE: no conv stack), keyed by geometry
# ----------------------------------------------------------------------------------------------
_post_engines: Dict[Tuple[int, int, int, int], Engine] = {}


def _post_engine(net_h: int, net_w: int, nb_class: int, device: int = 0) -> Engine:
    key = (int(net_h), int(net_w), int(nb_class), int(device))
    eng = _post_engines.get(key)
    if eng is None:
        eng = Engine(key[0], key[1], head=L.HEAD_NONE, nb_class=key[2], max_batch=1, device=device)
        _post_engines[key] = eng
    return eng


def _box_engine(n: int, nb_class: int) -> Engine:
    """A handle whose candidate capacity covers n boxes (do_nms / bbox_iou take arbitrary lists)."""
    size = 416
    while 3 * sum((size // s) ** 2 for s in (32, 16, 8)) < max(n, 2):
        size += 192
    return _post_engine(size, size, nb_class)


def _box_kind(boxes) -> str:
    """Arithmetic type the reference's type-generic bbox_iou / do_nms would run in for these coordinates under this NumPy (>= 2:
    Python scalars are weak): 'int' (pixel boxes after correct_yolo_boxes: exact integers, one double divide), 'f32' (np.float32
    coordinates, float32 operations) or 'f64' (Python floats / np.float64)."""
    has32 = has64 = hasf = False
    for b in boxes:
        for v in (b.xmin, b.ymin, b.xmax, b.ymax):
            if isinstance(v, (bool, np.bool_)):
                raise TypeError("boolean box coordinate")
            if isinstance(v, (int, np.integer)):
                continue
            if isinstance(v, np.float32):
                has32 = True
            elif isinstance(v, np.floating):
                has64 = True
            elif isinstance(v, float):
                hasf = True
            else:
                raise TypeError(f"unsupported box coordinate type {type(v)}")
    if has64:
        return "f64"
    if has32:
        return "f32"
    return "f64" if hasf else "int"


def _as_i32_boxes(boxes) -> np.ndarray:
    out = np.empty((len(boxes), 4), np.int32)
    for i, b in enumerate(boxes):
        for k, v in enumerate((b.xmin, b.ymin, b.xmax, b.ymax)):
            if not -(1 << 30) < int(v) < (1 << 30):
                raise OverflowError("box coordinate outside +-2^30")
            out[i, k] = int(v)
    return out


def _as_f64_boxes(boxes) -> np.ndarray:
    return np.array([[b.xmin, b.ymin, b.xmax, b.ymax] for b in boxes], np.float64).reshape(len(boxes), 4)


# ----------------------------------------------------------------------------------------------
# reference functions
# ----------------------------------------------------------------------------------------------
def bbox_iou(box1, box2):
    """``float(intersect) / union`` evaluated on the device (:183-194), in the arithmetic the coordinates' types select
    (integer pixel boxes: exact, one double divide; float boxes: every step a float operation, see fvy_bbox_iou_fp).  A zero
    union raises ZeroDivisionError for Python numbers exactly as the reference does; numpy scalars give nan / inf."""
    kind = _box_kind([box1, box2])
    coords = (box1.xmin, box1.xmax, box1.ymin, box1.ymax, box2.xmin, box2.xmax, box2.ymin, box2.ymax)
    python_only = all(isinstance(c, (int, float)) for c in coords)
    if kind == "int":
        ib = _as_i32_boxes([box1, box2])
        v = float(_box_engine(2, 1).bbox_iou(ib[0:1], ib[1:2])[0])
        if v != v and python_only:
            raise ZeroDivisionError("float division by zero")
        return v
    fb = _as_f64_boxes([box1, box2])
    v = _box_engine(2, 1).bbox_iou_fp(fb[0:1], fb[1:2], L.ARITH_F32 if kind == "f32" else L.ARITH_F64)[0]
    if python_only:
        if not np.isfinite(v):
            raise ZeroDivisionError("float division by zero")
        return float(v)
    return np.float32(v) if kind == "f32" else np.float64(v)


def decode_netout(netout, anchors, anchor_idx, obj_thresh, net_h, net_w, anchor_mask=None) -> List[BoundBox]:
    """One scale of one image -> BoundBox list in (row, col, anchor) order (:335-387).

    ``anchor_mask`` (3 bits, bit b = anchor b decoded) defaults to the fork's mask for ``anchor_idx`` (:354-362).  As in the
    reference, the caller's ``netout`` is modified IN PLACE - ``[..., :2]`` and ``[..., 4:]`` become their sigmoid (:343-344) -
    and every box's ``classes`` is a VIEW into it (:366), so do_nms' zeroing writes through.  (An argument that is not a writable
    C-contiguous float32 array cannot be updated in place; the views then point into a private copy.)"""
    raw = np.array(netout, dtype=np.float32, copy=True, order="C")          # the logits the device decodes
    gh, gw = raw.shape[:2]
    nb_class = raw.shape[-1] // 3 - 5
    if raw.ndim != 3 or raw.shape[-1] != 3 * (5 + nb_class) or nb_class < 1:
        raise ValueError(f"netout must be (grid_h, grid_w, 3*(5+nb_class)), got {raw.shape}")
    if anchor_idx not in (0, 1, 2):
        raise ValueError("anchor_idx must be 0, 1 or 2")
    eng = _post_engine(net_h, net_w, nb_class)
    if (gh, gw) != eng.grids[anchor_idx]:
        raise ValueError(f"scale {anchor_idx} of a {net_h}x{net_w} net has grid {eng.grids[anchor_idx]}, got {(gh, gw)}")
    if anchor_mask is None:
        anchor_mask = (0b010, 0b101, 0b010)[anchor_idx]
    anc = [0] * 18
    anc[6 * anchor_idx:6 * anchor_idx + 6] = [int(a) for a in anchors]
    pp = post_params(obj_thresh=obj_thresh, anchor_mask=(int(anchor_mask) & 7) << (3 * anchor_idx), anchors=anc, arith=DECODE_ARITH)
    outs = [np.zeros((1, g[0], g[1], raw.shape[-1]), np.float32) for g in eng.grids]
    outs[anchor_idx] = raw[None]
    d = eng.decode(outs, pp=pp, image_hw=None)
    n = int(d["counts"][0])
    # the in-place half, on the caller's array when it can be written
    inplace = isinstance(netout, np.ndarray) and netout.dtype == np.float32 and netout.flags["C_CONTIGUOUS"] and netout.flags["WRITEABLE"]
    view = (netout if inplace else raw.copy()).reshape(gh, gw, 3, 5 + nb_class)
    eng.netout_sigmoid(view)
    ftype = np.float64 if DECODE_ARITH == L.ARITH_F64 else np.float32
    first = int(3 * sum(g[0] * g[1] for g in eng.grids[:anchor_idx]))        # all-anchor numbering of fvy_det.cand
    boxes = []
    for i in range(n):
        x0, y0, x1, y1 = (ftype(v) for v in d["nbox"][0, i])
        cell, b = divmod(int(d["cand"][0, i]) - first, 3)
        row, col = divmod(cell, gw)
        boxes.append(BoundBox(x0, y0, x1, y1, view[row, col, b, 4], view[row, col, b, 5:], (anchors[2 * b + 0], anchors[2 * b + 1])))
    return boxes


def correct_yolo_boxes(boxes, image_h, image_w, net_h, net_w):
    """Letterbox inverse with int() truncation, in place (:389-404)."""
    if len(boxes) == 0:
        return
    nbox = np.array([[b.xmin, b.ymin, b.xmax, b.ymax] for b in boxes], np.float64)
    ib = _box_engine(len(boxes), 1).correct_boxes(nbox, image_h, image_w, net_h, net_w, DECODE_ARITH)
    for b, q in zip(boxes, ib):
        b.xmin, b.ymin, b.xmax, b.ymax = int(q[0]), int(q[1]), int(q[2]), int(q[3])


def correct_yolo_boxes_v2(boxes, image_size, net_h, net_w):
    """Same arithmetic with image_size = (image_h, image_w) (:406-424)."""
    correct_yolo_boxes(boxes, image_size[0], image_size[1], net_h, net_w)


def do_nms(boxes, nms_thresh):
    """Greedy per-class NMS, zeroing ``classes[c]`` of suppressed boxes in place (:426-444)."""
    if len(boxes) == 0:
        return
    nb_class = len(boxes[0].classes)
    _nms_inplace(boxes, nms_thresh, nb_class)


def do_nms_v2(boxes, nms_thresh):
    """Class-0-only variant used by FaceDetector.detect (:446-458)."""
    if len(boxes) == 0:
        return
    _nms_inplace(boxes, nms_thresh, 1)


def _nms_inplace(boxes, nms_thresh, nb_class):
    n = len(boxes)
    eng = _box_engine(n, nb_class)
    S = eng.cap
    cls = np.zeros((1, S, nb_class), np.float32)
    for i, b in enumerate(boxes):
        for c in range(nb_class):
            cls[0, i, c] = b.classes[c]
    kind = _box_kind(boxes)
    if kind == "int":          # the reference pipeline: pixel boxes after correct_yolo_boxes (:601-604)
        ib = np.zeros((1, S, 4), np.int32)
        ib[0, :n] = _as_i32_boxes(boxes)
        out, _, _ = eng.nms(ib, cls, np.array([n], np.int32), nms_thresh, want_kept=False)
    else:                      # float boxes: every IoU step a float operation of that type (fvy_nms_fp)
        fb = np.zeros((1, S, 4), np.float64)
        fb[0, :n] = _as_f64_boxes(boxes)
        out, _, _ = eng.nms_fp(fb, cls, np.array([n], np.int32), nms_thresh, L.ARITH_F32 if kind == "f32" else L.ARITH_F64, want_kept=False)
    for i, b in enumerate(boxes):
        for c in range(nb_class):
            if out[0, i, c] == 0 and b.classes[c] != 0:
                b.classes[c] = 0


# ----------------------------------------------------------------------------------------------
# model + weights
# ----------------------------------------------------------------------------------------------
class Yolov3Model:
    """What ``make_yolov3_model()`` returns: ``predict(x)`` = Keras ``Model.predict`` (:593).

    x: numpy (B, H, W, 3) float in [0,1], H and W multiples of 32 -> [ (B,H/32,W/32,C), (B,H/16,W/16,C),
    (B,H/8,W/8,C) ] float32.  Engines are created per input geometry on first use."""

    def __init__(self, nb_class=80, device=0, head=L.HEAD_YOLO3, tile_n_max=0):
        self.nb_class = nb_class
        self.device = device
        self.head = head
        self.tile_n_max = tile_n_max
        self._stream = None
        self._engines: Dict[Tuple[int, int, int], Engine] = {}

    @property
    def specs(self):
        return arch.table(self.head, self.nb_class)

    def count_params(self) -> int:
        return arch.n_params(self.specs)

    def set_weight_stream(self, stream: np.ndarray):
        stream = np.ascontiguousarray(stream, np.float32)
        if stream.size != self.count_params():
            raise ValueError(f"weight stream has {stream.size} floats, model needs {self.count_params()}")
        self._stream = stream
        for eng in self._engines.values():
            eng.load_weights(stream)

    def engine(self, batch: int, h: int, w: int) -> Engine:
        key = (h, w, batch)
        eng = self._engines.get(key)
        if eng is None:
            for (kh, kw, kb), e in self._engines.items():
                if kh == h and kw == w and kb >= batch:
                    return e
            if self._stream is None:   # Keras would run with its random initialisation (glorot-uniform, identity BN)
                self._stream = synth.darknet_stream(self.specs, 0, synth.INIT_KERAS_DEFAULT)
            eng = Engine(h, w, head=self.head, nb_class=self.nb_class, max_batch=batch, device=self.device, tile_n_max=self.tile_n_max)
            eng.load_weights(self._stream)
            self._engines[key] = eng
        return eng

    def predict(self, x, batch_size=None, verbose=0):
        x = np.ascontiguousarray(x)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float32)
        if x.ndim != 4 or x.shape[-1] != 3:
            raise ValueError("predict expects (B, H, W, 3)")
        b, h, w = x.shape[:3]
        outs = self.engine(b, h, w).forward(x)
        return outs if self.head == L.HEAD_YOLO3 else outs[0]

    def save(self, path):
        """Flat float32 blob in Darknet stream order (h5py is not available for Keras .h5)."""
        if self._stream is None:
            raise RuntimeError("no weights to save")
        with open(path, "wb") as f:
            f.write(synth.darknet_file_bytes(self._stream))


def make_yolov3_model(nb_class=80, device=0) -> Yolov3Model:
    """The reference hard-codes 255 = 3*(5+80) head channels (:278,294,308); ``nb_class=1`` gives the
    18-channel face heads of BASELINE.json."""
    return Yolov3Model(nb_class=nb_class, device=device)


class WeightReader:
    """Darknet ``yolov3.weights`` parser (:67-124): 3 x int32 header, 8- or 4-byte ``seen``, float32 stream."""

    def __init__(self, weight_file):
        with open(weight_file, "rb") as w_f:
            major, = struct.unpack("i", w_f.read(4))
            minor, = struct.unpack("i", w_f.read(4))
            revision, = struct.unpack("i", w_f.read(4))
            if (major * 10 + minor) >= 2 and major < 1000 and minor < 1000:
                w_f.read(8)
            else:
                w_f.read(4)
            binary = w_f.read()
        self.offset = 0
        self.all_weights = np.frombuffer(binary, dtype="float32")

    def read_bytes(self, size):
        self.offset = self.offset + size
        return self.all_weights[self.offset - size:self.offset]

    def load_weights(self, model: Yolov3Model):
        """Hands the stream to the model in the order the reference reads it (:91-119); BN folding and
        the (out,in,h,w) -> GEMM-operand repack happen in fvy_load_weights."""
        n = model.count_params()
        if self.all_weights.size < n:
            raise ValueError(f"weight file holds {self.all_weights.size} floats, model needs {n}")
        model.set_weight_stream(self.read_bytes(n))

    def reset(self):
        self.offset = 0


def _sigmoid(x):
    """Host helper kept for API parity (:180-181); the decode kernels evaluate it on the device."""
    return 1. / (1. + np.exp(-x))


def _interval_overlap(interval_a, interval_b):
    x1, x2 = interval_a
    x3, x4 = interval_b
    if x3 < x1:
        return 0 if x4 < x1 else min(x2, x4) - x1
    return 0 if x2 < x3 else min(x2, x4) - x3


__all__ = ["BoundBox", "WeightReader", "Yolov3Model", "make_yolov3_model", "decode_netout", "correct_yolo_boxes",
           "correct_yolo_boxes_v2", "do_nms", "do_nms_v2", "bbox_iou", "REF_ANCHORS"]
