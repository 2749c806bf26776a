"""Batch sharding for multi-GPU inference: one process per GPU, contiguous image slices, weights
replicated, NO data-path collective.  Replaces nothing in the reference's inference path
(it is strictly batch-1, face_detection.py:651-697); for training the reference uses Keras
``multi_gpu_model`` (face_detection.py:330,369), whose batch split along axis 0 this mirrors.
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
