"""ctypes binding of libfvy.so (the C ABI declared in include/fvy.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``tools/build.sh``.  There is no
CPU fallback: if the shared object is missing the import of this module raises, and every
compute entry point fails with FVY_E_CUDA when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FVY_LIB_PATH") or os.path.join(_HERE, "libfvy.so")   # FVY_LIB_PATH: A/B runs against another build

FVY_OK = 0
FVY_E_INVALID, FVY_E_CUDA, FVY_E_STATE, FVY_E_CAPACITY, FVY_E_RANGE = -1, -2, -3, -4, -5
HEAD_YOLO3, HEAD_FD6, HEAD_NONE = 0, 1, 2
F32, F64, U8 = 0, 1, 2
ARITH_F64, ARITH_F32 = 0, 1
ANCHOR_MASK_REFERENCE = 0x0AA   # src/space/yolov3_detect.py:354-362
ANCHOR_MASK_ALL = 0x1FF
CFG_NO_GRAPH, CFG_NO_CHAIN, CFG_NO_TILE_FLAGS, CFG_CHAIN_SCHED, CFG_NO_CHAIN_SCHED, CFG_NO_OVERLAP_POST, CFG_NO_FUSED_STEM = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40   # fvy_config.flags
CFG_NO_COMPACT = 0x80


class FvyConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("net_h", C.c_int), ("net_w", C.c_int), ("head", C.c_int), ("nb_class", C.c_int),
                ("bb_info_c_size", C.c_int), ("max_batch", C.c_int), ("max_cands", C.c_int), ("tile_n_max", C.c_int),
                ("flags", C.c_int)]


class FvyDet(C.Structure):
    _fields_ = [("xmin", C.c_int32), ("ymin", C.c_int32), ("xmax", C.c_int32), ("ymax", C.c_int32),
                ("objness", C.c_float), ("score", C.c_float), ("label", C.c_int32), ("cand", C.c_int32)]


class FvyPostParams(C.Structure):
    _fields_ = [("obj_thresh", C.c_double), ("nms_thresh", C.c_double), ("num_cands", C.c_int),
                ("anchor_mask", C.c_uint), ("anchors", C.c_int * 18), ("arith", C.c_int)]


# every symbol include/fvy.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = ["fvy_last_error", "fvy_version", "fvy_create", "fvy_destroy", "fvy_load_weights", "fvy_weight_count", "fvy_forward",
           "fvy_decode", "fvy_correct_boxes", "fvy_nms", "fvy_bbox_iou", "fvy_postprocess", "fvy_detect", "fvy_num_layers",
           "fvy_layer_info", "fvy_layer_output", "fvy_launch_count", "fvy_last_timing", "fvy_profile_layers", "fvy_run_layer", "fvy_timer_start", "fvy_timer_stop", "fvy_sync",
           "fvy_detect_async", "fvy_host_alloc", "fvy_host_free", "fvy_adam_step", "fvy_letterbox_u8", "fvy_staged_images", "fvy_read_staged",
           "fvy_bbox_iou_fp", "fvy_nms_fp", "fvy_netout_sigmoid", "fvy_timer_breakdown", "fvy_map_match",
           "fvy_bn_leaky_train_forward", "fvy_bn_leaky_train_backward", "fvy_conv_create", "fvy_conv_set_weights", "fvy_conv_run",
           "fvy_conv_wgrad_scratch_rows", "fvy_conv_wgrad"]

_lib = None


def load():
    """Load libfvy.so.  Raises (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the CUDA hot path)")
    L = C.CDLL(LIB_PATH)
    vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    H = C.c_void_p
    L.fvy_last_error.restype = C.c_char_p; L.fvy_last_error.argtypes = []
    L.fvy_version.restype = C.c_char_p; L.fvy_version.argtypes = []
    L.fvy_create.restype = C.c_int; L.fvy_create.argtypes = [C.POINTER(FvyConfig), C.POINTER(H)]
    L.fvy_destroy.restype = None; L.fvy_destroy.argtypes = [H]
    L.fvy_load_weights.restype = C.c_int; L.fvy_load_weights.argtypes = [H, vp, C.c_size_t]
    L.fvy_weight_count.restype = C.c_longlong; L.fvy_weight_count.argtypes = [H]
    L.fvy_forward.restype = C.c_int; L.fvy_forward.argtypes = [H, vp, C.c_int, C.c_int, vp, vp, vp]
    L.fvy_decode.restype = C.c_int
    L.fvy_decode.argtypes = [H, vp, vp, vp, C.c_int, C.POINTER(FvyPostParams), vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.fvy_correct_boxes.restype = C.c_int
    L.fvy_correct_boxes.argtypes = [H, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.fvy_nms.restype = C.c_int
    L.fvy_nms.argtypes = [H, vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, vp, vp]
    L.fvy_bbox_iou.restype = C.c_int; L.fvy_bbox_iou.argtypes = [H, vp, vp, C.c_int, vp]
    L.fvy_bbox_iou_fp.restype = C.c_int; L.fvy_bbox_iou_fp.argtypes = [H, vp, vp, C.c_int, C.c_int, vp]
    L.fvy_nms_fp.restype = C.c_int
    L.fvy_nms_fp.argtypes = [H, vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, vp, vp]
    L.fvy_netout_sigmoid.restype = C.c_int; L.fvy_netout_sigmoid.argtypes = [H, vp, C.c_longlong, C.c_int]
    L.fvy_map_match.restype = C.c_int; L.fvy_map_match.argtypes = [H, vp, vp, vp, vp, C.c_int, vp, vp]
    L.fvy_postprocess.restype = C.c_int
    L.fvy_postprocess.argtypes = [H, vp, vp, vp, C.c_int, C.POINTER(FvyPostParams), vp, C.c_int, vp, vp]
    for name in ("fvy_detect", "fvy_detect_async"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [H, vp, C.c_int, C.c_int, C.POINTER(FvyPostParams), vp, C.c_int, vp, vp]
    L.fvy_letterbox_u8.restype = C.c_int
    L.fvy_letterbox_u8.argtypes = [H, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fvy_staged_images.restype = C.c_void_p; L.fvy_staged_images.argtypes = [H]
    L.fvy_read_staged.restype = C.c_int; L.fvy_read_staged.argtypes = [H, C.c_int, vp]
    L.fvy_num_layers.restype = C.c_int; L.fvy_num_layers.argtypes = [H]
    L.fvy_layer_info.restype = C.c_int; L.fvy_layer_info.argtypes = [H, C.c_int, ip]
    L.fvy_layer_output.restype = C.c_int; L.fvy_layer_output.argtypes = [H, C.c_int, C.c_int, vp]
    L.fvy_launch_count.restype = C.c_longlong; L.fvy_launch_count.argtypes = [H]
    L.fvy_last_timing.restype = C.c_int; L.fvy_last_timing.argtypes = [H, fp, fp]
    L.fvy_profile_layers.restype = C.c_int; L.fvy_profile_layers.argtypes = [H, C.c_int, C.c_int, vp]
    L.fvy_run_layer.restype = C.c_int; L.fvy_run_layer.argtypes = [H, C.c_int, C.c_int, C.c_int, fp]
    L.fvy_timer_start.restype = C.c_int; L.fvy_timer_start.argtypes = [H]
    L.fvy_timer_stop.restype = C.c_int; L.fvy_timer_stop.argtypes = [H, fp]
    L.fvy_sync.restype = C.c_int; L.fvy_sync.argtypes = [H]
    L.fvy_timer_breakdown.restype = C.c_int; L.fvy_timer_breakdown.argtypes = [H, fp, fp, ip]
    L.fvy_host_alloc.restype = C.c_void_p; L.fvy_host_alloc.argtypes = [C.c_size_t]
    L.fvy_host_free.restype = None; L.fvy_host_free.argtypes = [C.c_void_p]
    L.fvy_bn_leaky_train_forward.restype = C.c_int
    L.fvy_bn_leaky_train_forward.argtypes = [vp, C.c_longlong, C.c_int, vp, vp, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp]
    L.fvy_bn_leaky_train_backward.restype = C.c_int
    L.fvy_conv_create.restype = C.c_int
    L.fvy_conv_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.fvy_conv_set_weights.restype = C.c_int
    L.fvy_conv_set_weights.argtypes = [vp, vp, C.c_int, vp]
    L.fvy_conv_run.restype = C.c_int
    L.fvy_conv_run.argtypes = [vp, vp, C.c_int, vp, vp]
    L.fvy_conv_wgrad_scratch_rows.restype = C.c_longlong
    L.fvy_conv_wgrad_scratch_rows.argtypes = [C.c_int, C.c_int, C.c_int]
    L.fvy_conv_wgrad.restype = C.c_int
    L.fvy_conv_wgrad.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.fvy_bn_leaky_train_backward.argtypes = [vp, vp, C.c_longlong, C.c_int, vp, vp, vp, vp, C.c_float, vp, vp, vp, vp, vp]
    L.fvy_adam_step.restype = C.c_int
    L.fvy_adam_step.argtypes = [vp, vp, vp, vp, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, vp]
    _lib = L
    return L


class FvyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fvy error {code}: {msg}")
        self.code = code


def check(code: int):
    """Map an fvy_status to a Python exception (SURVEY 8b error convention)."""
    if code == FVY_OK:
        return
    msg = load().fvy_last_error().decode("utf-8", "replace")
    if code == FVY_E_INVALID:
        raise ValueError(f"fvy: {msg}")
    if code == FVY_E_RANGE:
        raise OverflowError(f"fvy: {msg}")
    raise FvyError(code, msg)
