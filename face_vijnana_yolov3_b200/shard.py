"""Batch sharding for multi-GPU inference: contiguous image slices, weights replicated, NO data-path collective.
Two drivers: one process per GPU (``torchrun bench.py``: ``shard_bounds`` over the ranks) and one process for all GPUs
(``ShardedDetector``: one handle and one host thread per device).  The reference's inference path is strictly batch-1
(face_detection.py:651-697); for training it uses Keras ``multi_gpu_model`` (face_detection.py:330,369), whose batch split
along axis 0 this mirrors.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_bounds(global_batch: int, world_size: int) -> List[Tuple[int, int]]:
    """Image i goes to rank floor(i * world_size / global_batch): contiguous, sizes differ by <= 1."""
    if world_size < 1 or global_batch < 0:
        raise ValueError("bad shard arguments")
    bounds = []
    for r in range(world_size):
        lo = -(-r * global_batch // world_size)          # ceil(r*B/N)
        hi = -(-(r + 1) * global_batch // world_size)
        bounds.append((lo, hi))
    return bounds


def owner_of(image: int, global_batch: int, world_size: int) -> int:
    return image * world_size // global_batch


def gather_in_image_order(per_rank_results: List[list]) -> list:
    """Concatenate per-rank result lists (rank order == image order because slices are contiguous)."""
    out = []
    for r in per_rank_results:
        out.extend(r)
    return out


class ShardedDetector:
    """One batch, N devices, ONE call: the in-process replacement of the reference's ``multi_gpu_model`` call site
    (face_detection.py:330, :369) for inference.  One handle (Engine) per device, each driven by its own host thread (ctypes
    releases the GIL inside the C ABI); image i goes to device ``owner_of(i, B, N)``; the per-device results are written straight
    into one (B, max_out) record array, i.e. in image order.  No collective: every image's detections depend on that image alone,
    so they are bit-identical to the single-device result (tests/test_gpu_round2.py::test_sharded_detect_identity).

    ``devices`` may name a device more than once only for testing; such handles run one after the other (two persistent conv
    grids cannot share one GPU's SMs at the same time).
    """

    def __init__(self, devices, net_h=416, net_w=416, head=None, nb_class=1, max_batch_per_device=40, **engine_kw):
        from . import _lib as L
        from .engine import Engine
        if not devices:
            raise ValueError("ShardedDetector needs at least one device")
        self.devices = [int(d) for d in devices]
        self.head = L.HEAD_YOLO3 if head is None else head
        self.engines = [Engine(net_h, net_w, head=self.head, nb_class=nb_class, max_batch=max_batch_per_device, device=d, **engine_kw)
                        for d in self.devices]
        self.max_batch = max_batch_per_device * len(self.devices)
        self.cap = self.engines[0].cap
        import threading
        self._dev_locks = {d: threading.Lock() for d in set(self.devices)}

    def load_weights(self, stream) -> None:
        for e in self.engines:
            e.load_weights(stream)

    def close(self) -> None:
        for e in self.engines:
            e.close()

    def detect(self, images, pp=None, image_hw=None, max_out=None):
        """images (B, H, W, 3) host array (float32 / float64 / uint8) -> (dets (B, max_out), counts (B,)) in image order."""
        import threading
        import numpy as np
        from .engine import DET_DTYPE
        B = int(images.shape[0])
        n = len(self.engines)
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds {self.max_batch} (max_batch_per_device x devices)")
        max_out = int(max_out or self.cap)
        dets = np.zeros((B, max_out), DET_DTYPE)
        counts = np.zeros(B, np.int32)
        hw = None if image_hw is None else np.ascontiguousarray(image_hw, np.int32).reshape(B, 2)
        errors = [None] * n

        def work(r, lo, hi):
            try:
                with self._dev_locks[self.devices[r]]:
                    d, c = self.engines[r].detect(np.ascontiguousarray(images[lo:hi]), pp=pp, image_hw=None if hw is None else hw[lo:hi],
                                                  max_out=max_out)
                dets[lo:hi] = d
                counts[lo:hi] = c
            except BaseException as exc:      # re-raised on the calling thread
                errors[r] = exc

        threads = []
        for r, (lo, hi) in enumerate(shard_bounds(B, n)):
            if hi > lo:
                if hi - lo > self.engines[r].max_batch:
                    raise ValueError("shard larger than max_batch_per_device")
                t = threading.Thread(target=work, args=(r, lo, hi), daemon=True)
                t.start()
                threads.append(t)
        for t in threads:
            t.join()
        for exc in errors:
            if exc is not None:
                raise exc
        return dets, counts
