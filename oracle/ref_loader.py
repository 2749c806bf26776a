"""ORACLE — test infrastructure only.  Import the REAL reference post-processing code.

``/root/reference/src/space/yolov3_detect.py`` imports keras / skimage at module scope
(:32-39); neither is installed.  Registering empty stub modules first lets the file execute, and
its genuine ``decode_netout / correct_yolo_boxes(_v2) / do_nms(_v2) / bbox_iou / BoundBox``
(:126-458) run unmodified under this container's NumPy (>= 2: float32 scalar arithmetic, SURVEY
App. B-4).  ``face_detection.py`` is loaded the same way for ``FaceDetector.detect`` (:885-949).

Exists only in the build container: returns ``None`` when /root/reference is absent (GPU box).
"""
from __future__ import annotations

import os
import sys
import types

REF_SPACE = "/root/reference/src/space"


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Any()


def _stub(name, attrs=()):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__stub__"] = True
        sys.modules[name] = mod
    for a in attrs:
        if not hasattr(mod, a):
            setattr(mod, a, _Any)
    return mod


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SPACE, "yolov3_detect.py"))


_cache = {}


def load_yolov3_detect():
    """Return the reference ``yolov3_detect`` module, or None if the reference is absent."""
    if not available():
        return None
    if "y" in _cache:
        return _cache["y"]
    _stub("skimage"); _stub("skimage.io", ("imread", "imsave")); _stub("skimage.transform", ("resize",))
    _stub("skimage.draw", ("polygon_perimeter", "set_color"))
    _stub("keras"); _stub("keras.layers", ("Conv2D", "Input", "BatchNormalization", "LeakyReLU", "ZeroPadding2D",
                                           "UpSampling2D", "Lambda", "Concatenate", "Dense"))
    _stub("keras.layers.merge", ("add", "concatenate")); _stub("keras.models", ("Model", "load_model"))
    if REF_SPACE not in sys.path:
        sys.path.insert(0, REF_SPACE)
    import importlib
    saved_argv = sys.argv
    try:
        sys.argv = [saved_argv[0]]
        mod = importlib.import_module("yolov3_detect")
    finally:
        sys.argv = saved_argv
    _cache["y"] = mod
    return mod


def load_face_detection():
    """Return the reference ``face_detection`` module (for FaceDetector.detect), or None."""
    y = load_yolov3_detect()
    if y is None:
        return None
    if "f" in _cache:
        return _cache["f"]
    _stub("keras.utils", ("multi_gpu_model",)); _stub("keras.optimizers", ("Adam",))
    ku = _stub("keras.utils.data_utils")
    ku.Sequence = object
    _stub("keras.backend")
    sys.modules["keras"].optimizers = sys.modules["keras.optimizers"]
    sys.modules["keras"].backend = sys.modules["keras.backend"]
    import importlib
    try:
        mod = importlib.import_module("face_detection")
    except Exception as e:  # pragma: no cover - depends on what else the file imports
        _cache["f"] = None
        raise RuntimeError(f"reference face_detection not importable with stubs: {e}")
    _cache["f"] = mod
    return mod


def make_ref_face_detector(hps, image_size=416, predict_fn=None):
    """Build the reference ``FaceDetector`` without Keras: ``__new__`` + hand-set attributes
    (face_detection.py:320-325) + a fake ``.model.predict`` returning the (1,13,13,6) map."""
    fdm = load_face_detection()
    fd = fdm.FaceDetector.__new__(fdm.FaceDetector)
    fd.hps = dict(hps)
    fd.nn_arch = {"image_size": image_size, "bb_info_c_size": 6}
    fd.cell_image_size = image_size // fdm.FaceDetector.CELL_SIZE
    fd.model = types.SimpleNamespace(predict=predict_fn)
    return fd


def load_evaluate():
    """Return the reference ``evaluate`` module (for ``cal_mAP_fd``, evaluate.py:27-127), or None.  h5py is stubbed (only ``main``
    writes with it).  One pandas shim is installed: the reference does ``sol_df.iat[:, 6] = -1.0`` (evaluate.py:31, :36), which the
    pandas of its day accepted and pandas >= 1 rejects (``iat`` takes scalar positions); the shim routes a slice key of
    ``.iat.__setitem__`` to ``.iloc`` - the assignment the line means.  Nothing of the reference's source is altered."""
    y = load_yolov3_detect()
    if y is None:
        return None
    if "e" in _cache:
        return _cache["e"]
    _stub("h5py")
    import pandas.core.indexing as pci
    if not getattr(pci._iAtIndexer, "__fvy_shim__", False):
        orig = pci._iAtIndexer.__setitem__

        def setitem(self, key, value):
            if isinstance(key, tuple) and any(isinstance(k, slice) for k in key):
                self.obj.iloc[key] = value
                return
            return orig(self, key, value)
        pci._iAtIndexer.__setitem__ = setitem
        pci._iAtIndexer.__fvy_shim__ = True
    import importlib
    mod = importlib.import_module("evaluate")
    mod.DEBUG = False
    _cache["e"] = mod
    return mod
