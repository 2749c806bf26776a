"""Python face of one fvy_handle: the Darknet-53/YOLOv3 conv stack + decode + NMS on one B200.

Thin host code over the C ABI (include/fvy.h).  numpy arrays are passed as host pointers, torch
CUDA tensors (anything with ``data_ptr()`` and ``is_cuda``) as device pointers.  The reference
interfaces replaced are cited per method (paths under /root/reference/src/space/).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from . import arch

REF_ANCHORS = (116, 90, 156, 198, 373, 326, 30, 61, 62, 45, 59, 119, 10, 13, 16, 30, 33, 23)   # yolov3_detect.py:558-560

DET_DTYPE = np.dtype([("xmin", "<i4"), ("ymin", "<i4"), ("xmax", "<i4"), ("ymax", "<i4"), ("objness", "<f4"),
                      ("score", "<f4"), ("label", "<i4"), ("cand", "<i4")])
assert DET_DTYPE.itemsize == C.sizeof(L.FvyDet) == 32


class StagedImages:
    """Device-resident float32 (B, H, W, 3) view of the handle's staged batch (fvy_staged_images)."""

    def __init__(self, ptr, shape):
        self._ptr, self.shape, self.dtype = int(ptr), tuple(shape), "float32"

    def data_ptr(self):
        return self._ptr

    def is_contiguous(self):
        return True


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):   # torch tensor (host or device)
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(x.data_ptr())
    raise TypeError(f"unsupported buffer type {type(x)}")


def post_params(obj_thresh=0.5, nms_thresh=0.45, num_cands=0, anchor_mask=L.ANCHOR_MASK_REFERENCE, anchors=REF_ANCHORS,
                arith=L.ARITH_F64) -> L.FvyPostParams:
    p = L.FvyPostParams()
    p.obj_thresh = float(obj_thresh); p.nms_thresh = float(nms_thresh); p.num_cands = int(num_cands)
    p.anchor_mask = int(anchor_mask); p.arith = int(arith)
    flat = [int(v) for v in np.asarray(anchors).ravel()]
    if len(flat) != 18:
        raise ValueError("anchors must hold 9 (w,h) pairs")
    for i, v in enumerate(flat):
        p.anchors[i] = v
    return p


class Engine:
    """One handle = one device + one stream.  Not re-entrant (SURVEY 8b threading)."""

    def __init__(self, net_h=416, net_w=416, head=L.HEAD_YOLO3, nb_class=1, max_batch=1, device=0, max_cands=0,
                 bb_info_c_size=6, tile_n_max=0, flags=0):
        self.lib = L.load()
        cfg = L.FvyConfig(device=device, net_h=net_h, net_w=net_w, head=head, nb_class=nb_class, bb_info_c_size=bb_info_c_size,
                          max_batch=max_batch, max_cands=max_cands, tile_n_max=tile_n_max, flags=int(flags))
        self.cfg = cfg
        self._h = C.c_void_p()
        L.check(self.lib.fvy_create(C.byref(cfg), C.byref(self._h)))
        self.head = head
        self.nb_class = nb_class if head != L.HEAD_FD6 else 1
        self.max_batch = max_batch
        self.net_h, self.net_w = net_h, net_w
        if head == L.HEAD_FD6:
            self.head_c = bb_info_c_size
            self.grids = [(net_h // 32, net_w // 32)]
            self.cap = max_cands or self.grids[0][0] * self.grids[0][1]
        else:
            self.head_c = 3 * (5 + nb_class)
            self.grids = [(net_h // s, net_w // s) for s in (32, 16, 8)]
            self.cap = max_cands or sum(3 * a * b for a, b in self.grids)

    # ---------------------------------------------------------------- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.fvy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- weights (WeightReader.load_weights, yolov3_detect.py:90-121)
    @property
    def weight_count(self) -> int:
        return int(self.lib.fvy_weight_count(self._h))

    def load_weights(self, stream: np.ndarray):
        stream = np.ascontiguousarray(stream, dtype=np.float32)
        L.check(self.lib.fvy_load_weights(self._h, _ptr(stream), stream.size))

    # ---------------------------------------------------------------- forward (Model.predict, yolov3_detect.py:593 / face_detection.py:899)
    def forward(self, images, outs: Optional[Sequence] = None, want_outputs=True) -> Optional[List[np.ndarray]]:
        dtype, batch = self._image_args(images)
        if outs is None and want_outputs:
            outs = [np.empty((batch, gh, gw, self.head_c), np.float32) for gh, gw in self.grids]
        o = list(outs) if outs is not None else []
        o += [None] * (3 - len(o))
        L.check(self.lib.fvy_forward(self._h, _ptr(images), dtype, batch, _ptr(o[0]), _ptr(o[1]), _ptr(o[2])))
        return list(outs) if outs is not None else None

    def _image_args(self, images):
        shape = tuple(images.shape)
        if len(shape) != 4 or shape[1:] != (self.net_h, self.net_w, 3):
            raise ValueError(f"images must be (B,{self.net_h},{self.net_w},3) NHWC, got {shape}")
        if isinstance(images, np.ndarray):
            if images.dtype == np.float32:
                dtype = L.F32
            elif images.dtype == np.float64:
                dtype = L.F64
            elif images.dtype == np.uint8:
                dtype = L.U8
            else:
                raise TypeError("images must be float32 / float64 in [0,1] or uint8 in 0..255")
        else:
            s = str(images.dtype)
            dtype = L.F32 if s.endswith("float32") else L.F64 if s.endswith("float64") else L.U8 if s.endswith("uint8") else None
            if dtype is None:
                raise TypeError("images must be float32 / float64 in [0,1] or uint8 in 0..255")
        return dtype, int(shape[0])

    # ---------------------------------------------------------------- decode_netout + correct_yolo_boxes (yolov3_detect.py:335-404)
    def decode(self, outs=None, batch=None, pp: Optional[L.FvyPostParams] = None, image_hw=None, want_nbox=True):
        pp = pp or post_params()
        if outs is not None:
            batch = int(outs[0].shape[0])
        if batch is None:
            raise ValueError("batch is required when decoding the resident logits")
        cap = self.cap
        nbox = np.empty((batch, cap, 4), np.float64) if want_nbox and self.head != L.HEAD_FD6 else None
        hw = None if image_hw is None else np.ascontiguousarray(image_hw, np.int32).reshape(batch, 2)
        ibox = np.empty((batch, cap, 4), np.int32) if (hw is not None or self.head == L.HEAD_FD6) else None
        obj = np.empty((batch, cap), np.float32)
        cls = np.empty((batch, cap, self.nb_class), np.float32)
        cand = np.empty((batch, cap), np.int32)
        counts = np.empty(batch, np.int32)
        o = list(outs) if outs is not None else []
        o = [np.ascontiguousarray(x, np.float32) if isinstance(x, np.ndarray) else x for x in o] + [None] * (3 - len(o))
        L.check(self.lib.fvy_decode(self._h, _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), batch, C.byref(pp), _ptr(hw), cap, _ptr(nbox),
                                    _ptr(ibox), _ptr(obj), _ptr(cls), _ptr(cand), _ptr(counts)))
        return dict(nbox=nbox, ibox=ibox, objness=obj, classes=cls, cand=cand, counts=counts)

    def correct_boxes(self, nbox, image_h, image_w, net_h, net_w, arith=L.ARITH_F64) -> np.ndarray:
        nbox = np.ascontiguousarray(nbox, np.float64).reshape(-1, 4)
        ibox = np.empty((nbox.shape[0], 4), np.int32)
        L.check(self.lib.fvy_correct_boxes(self._h, _ptr(nbox), nbox.shape[0], int(image_h), int(image_w), int(net_h), int(net_w),
                                           int(arith), _ptr(ibox)))
        return ibox

    # ---------------------------------------------------------------- do_nms / do_nms_v2 (yolov3_detect.py:426-458)
    def nms(self, ibox, classes, counts, nms_thresh, want_kept=True):
        """ibox (B,S,4) int32, classes (B,S,nc) float32 (modified copy returned), counts (B,)."""
        ibox = np.ascontiguousarray(ibox, np.int32)
        classes = np.array(classes, np.float32, copy=True, order="C")
        if classes.ndim == 2:
            classes = classes[:, :, None]
        B, S, nc = classes.shape
        counts = np.ascontiguousarray(counts, np.int32)
        kept = np.empty((B, S), np.int32) if want_kept else None
        kc = np.empty(B, np.int32) if want_kept else None
        L.check(self.lib.fvy_nms(self._h, _ptr(ibox), _ptr(counts), B, S, nc, float(nms_thresh), _ptr(classes), _ptr(kept), _ptr(kc)))
        return classes, kept, kc

    def bbox_iou(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a, np.int32).reshape(-1, 4); b = np.ascontiguousarray(b, np.int32).reshape(-1, 4)
        out = np.empty(a.shape[0], np.float64)
        L.check(self.lib.fvy_bbox_iou(self._h, _ptr(a), _ptr(b), a.shape[0], _ptr(out)))
        return out

    def bbox_iou_fp(self, a, b, arith=L.ARITH_F64) -> np.ndarray:
        """bbox_iou on float boxes (yolov3_detect.py:183-194 before correct_yolo_boxes / evaluate.py:69); see fvy_bbox_iou_fp."""
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 4); b = np.ascontiguousarray(b, np.float64).reshape(-1, 4)
        out = np.empty(a.shape[0], np.float64)
        L.check(self.lib.fvy_bbox_iou_fp(self._h, _ptr(a), _ptr(b), a.shape[0], int(arith), _ptr(out)))
        return out

    def nms_fp(self, box, classes, counts, nms_thresh, arith=L.ARITH_F64, want_kept=True):
        """do_nms on float boxes: box (B,S,4) float64, classes (B,S,nc) float32 (modified copy returned), counts (B,)."""
        box = np.ascontiguousarray(box, np.float64)
        classes = np.array(classes, np.float32, copy=True, order="C")
        if classes.ndim == 2:
            classes = classes[:, :, None]
        B, S, nc = classes.shape
        counts = np.ascontiguousarray(counts, np.int32)
        kept = np.empty((B, S), np.int32) if want_kept else None
        kc = np.empty(B, np.int32) if want_kept else None
        L.check(self.lib.fvy_nms_fp(self._h, _ptr(box), _ptr(counts), B, S, nc, float(nms_thresh), int(arith), _ptr(classes), _ptr(kept), _ptr(kc)))
        return classes, kept, kc

    def netout_sigmoid(self, netout4: np.ndarray) -> None:
        """In place: netout[..., :2] and netout[..., 4:] -> sigmoid, as decode_netout does to its argument (yolov3_detect.py:343-344)."""
        if netout4.dtype != np.float32 or not netout4.flags["C_CONTIGUOUS"] or not netout4.flags["WRITEABLE"] or netout4.ndim != 4:
            raise ValueError("netout_sigmoid expects a writable C-contiguous float32 (gh, gw, 3, 5+nb_class) array")
        gh, gw, nb, ch = netout4.shape
        L.check(self.lib.fvy_netout_sigmoid(self._h, _ptr(netout4), gh * gw * nb, ch - 5))

    def map_match(self, gt_box, gt_off, det_box, det_off):
        """Matching half of evaluate.cal_mAP_fd (evaluate.py:41-100): -> (det_iou[n_det] float64, img_any[n_img] int32); see fvy_map_match."""
        gt_box = np.ascontiguousarray(gt_box, np.float64).reshape(-1, 4); det_box = np.ascontiguousarray(det_box, np.float64).reshape(-1, 4)
        gt_off = np.ascontiguousarray(gt_off, np.int32); det_off = np.ascontiguousarray(det_off, np.int32)
        n_img = gt_off.size - 1
        if det_off.size != n_img + 1 or gt_off[-1] != gt_box.shape[0] or det_off[-1] != det_box.shape[0]:
            raise ValueError("offset arrays do not describe the box arrays")
        det_iou = np.full(det_box.shape[0], -1.0, np.float64)
        img_any = np.zeros(max(n_img, 0), np.int32)
        L.check(self.lib.fvy_map_match(self._h, _ptr(gt_box), _ptr(gt_off), _ptr(det_box), _ptr(det_off), n_img, _ptr(det_iou), _ptr(img_any)))
        return det_iou, img_any

    # ---------------------------------------------------------------- whole path
    def postprocess(self, outs=None, batch=None, pp=None, image_hw=None, max_out=None):
        pp = pp or post_params()
        if outs is not None:
            batch = int(outs[0].shape[0])
        max_out = max_out or self.cap
        dets = np.empty((batch, max_out), DET_DTYPE)
        counts = np.empty(batch, np.int32)
        hw = None if image_hw is None else np.ascontiguousarray(image_hw, np.int32).reshape(batch, 2)
        o = list(outs) if outs is not None else []
        o = [np.ascontiguousarray(x, np.float32) if isinstance(x, np.ndarray) else x for x in o] + [None] * (3 - len(o))
        L.check(self.lib.fvy_postprocess(self._h, _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), batch, C.byref(pp), _ptr(hw), max_out,
                                         _ptr(dets), _ptr(counts)))
        return dets, counts

    def detect(self, images, pp=None, image_hw=None, max_out=None, dets=None, counts=None, sync=True):
        """forward + decode + NMS (yolov3_detect._main_ :593-604 / FaceDetector.detect :899-947)."""
        pp = pp or post_params()
        dtype, batch = self._image_args(images)
        max_out = max_out or self.cap
        if dets is None:
            dets = np.empty((batch, max_out), DET_DTYPE)
        if counts is None:
            counts = np.empty(batch, np.int32)
        if image_hw is None and self.head != L.HEAD_FD6:
            image_hw = np.tile(np.array([self.net_h, self.net_w], np.int32), (batch, 1))
        hw = None if image_hw is None else (image_hw if not isinstance(image_hw, (list, tuple)) else np.asarray(image_hw, np.int32))
        if isinstance(hw, np.ndarray):
            hw = np.ascontiguousarray(hw, np.int32).reshape(batch, 2)
        fn = self.lib.fvy_detect if sync else self.lib.fvy_detect_async
        L.check(fn(self._h, _ptr(images), dtype, batch, C.byref(pp), _ptr(hw), max_out, _ptr(dets), _ptr(counts)))
        self._keep = (images, hw, dets, counts, pp)   # keep buffers alive for async use
        return dets, counts

    # ---------------------------------------------------------------- FaceDetector.evaluate / test pre-processing (face_detection.py:657-690)
    def letterbox(self, image_u8: np.ndarray, index: int, w_p: int, h_p: int, pad_t: int, pad_l: int) -> None:
        """image/255 -> cv.resize(INTER_CUBIC) to (w_p, h_p) -> zero border, on the GPU, into slot ``index`` of the staged batch."""
        if image_u8.dtype != np.uint8 or image_u8.ndim != 3 or image_u8.shape[2] != 3:
            raise TypeError("letterbox expects a uint8 (H, W, 3) image")
        image_u8 = np.ascontiguousarray(image_u8)
        L.check(self.lib.fvy_letterbox_u8(self._h, _ptr(image_u8), image_u8.shape[0], image_u8.shape[1], int(w_p), int(h_p), int(pad_t), int(pad_l), int(index)))

    def staged(self, batch: int) -> "StagedImages":
        """The first ``batch`` images of the staged (letterboxed) batch as an ``images`` argument of detect / forward."""
        if not 1 <= batch <= self.max_batch:
            raise ValueError("batch out of range")
        ptr = self.lib.fvy_staged_images(self._h)
        if not ptr:
            raise L.FvyError(-2, "the staged batch could not be allocated")
        return StagedImages(ptr, (batch, self.net_h, self.net_w, 3))

    def staged_to_host(self, batch: int) -> np.ndarray:
        """Host copy of the staged batch (tests / inspection)."""
        out = np.empty((batch, self.net_h, self.net_w, 3), np.float32)
        L.check(self.lib.fvy_read_staged(self._h, int(batch), _ptr(out)))
        return out

    def timer_start(self):
        L.check(self.lib.fvy_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        L.check(self.lib.fvy_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def timer_breakdown(self):
        """(mean forward ms, mean post-processing ms, calls) of the detect calls since timer_start, from CUDA events in the timed region."""
        a, b, n = C.c_float(), C.c_float(), C.c_int()
        L.check(self.lib.fvy_timer_breakdown(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def sync(self):
        L.check(self.lib.fvy_sync(self._h))

    # ---------------------------------------------------------------- introspection
    def layer_infos(self):
        n = self.lib.fvy_num_layers(self._h)
        keys = ("idx", "cin", "cout", "k", "stride", "H", "W", "tile_n", "tile_k", "stages", "grid", "tiles")
        out = []
        for i in range(n):
            buf = (C.c_int * 12)()
            L.check(self.lib.fvy_layer_info(self._h, i, buf))
            out.append(dict(zip(keys, list(buf))))
        return out

    def layer_output(self, layer: int, batch: int) -> np.ndarray:
        info = self.layer_infos()[layer]
        dst = np.empty((batch, info["H"], info["W"], info["cout"]), np.float32)
        L.check(self.lib.fvy_layer_output(self._h, layer, batch, _ptr(dst)))
        return dst

    @property
    def launch_count(self) -> int:
        return int(self.lib.fvy_launch_count(self._h))

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        L.check(self.lib.fvy_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile_layers(self, batch: int, iters: int = 5) -> np.ndarray:
        ms = np.zeros(self.lib.fvy_num_layers(self._h), np.float32)
        L.check(self.lib.fvy_profile_layers(self._h, batch, iters, _ptr(ms)))
        return ms

    def run_layer(self, layer: int, batch: int, iters: int = 3) -> float:
        ms = C.c_float()
        L.check(self.lib.fvy_run_layer(self._h, layer, batch, iters, C.byref(ms)))
        return ms.value

    def macs_per_image(self) -> int:
        return arch.macs(arch.table(self.head, self.nb_class), self.net_h, self.net_w)
