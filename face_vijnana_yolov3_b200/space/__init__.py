"""Drop-ins for the hot-path modules of the reference's ``src/space`` package; re-exports what its ``__init__`` does (:1-3)."""
from . import yolov3_detect                      # noqa: F401
from .yolov3_detect import BoundBox, bbox_iou    # noqa: F401
