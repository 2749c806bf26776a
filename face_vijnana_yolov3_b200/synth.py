"""Seeded synthetic inputs: Darknet-format weight streams, images, head logits.

There is no network for ``yolov3.weights`` / ``face_detector.h5``, so every config in
BASELINE.json runs on random-init weights.  The stream layout is the one the
reference's ``WeightReader.load_weights`` consumes (``src/space/yolov3_detect.py:91-119``):
for each conv in ascending layer index, ``beta, gamma, mean, var`` (bn layers) or ``bias``
(linear heads) followed by the kernel in (Cout, Cin, kh, kw) order, all float32.
Not on the hot path; used by tests, bench.py and smoke().
"""
from __future__ import annotations

import numpy as np

from . import arch

INIT_KERAS_DEFAULT = "keras_default"   # glorot-uniform kernels, zero bias, BN gamma=1 beta=0 mean=0 var=1
INIT_BN_EXERCISING = "bn_exercising"   # kernels N(0,1/fan_in), gamma,var~U(.75,1.25), beta,mean~N(0,.1), head bias N(0,.1)


def darknet_stream(specs, seed: int = 0, init: str = INIT_BN_EXERCISING) -> np.ndarray:
    """Return the float32 weight stream (header excluded) for ``specs``."""
    parts = []
    for c in specs:
        rng = np.random.default_rng([seed, c.idx])
        fan_in = c.k * c.k * c.cin
        fan_out = c.k * c.k * c.cout
        if init == INIT_KERAS_DEFAULT:
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            kern = rng.uniform(-lim, lim, size=(c.cout, c.cin, c.k, c.k)).astype(np.float32)
            if c.bn:
                beta = np.zeros(c.cout, np.float32); gamma = np.ones(c.cout, np.float32)
                mean = np.zeros(c.cout, np.float32); var = np.ones(c.cout, np.float32)
            else:
                bias = np.zeros(c.cout, np.float32)
        elif init == INIT_BN_EXERCISING:
            kern = (rng.standard_normal((c.cout, c.cin, c.k, c.k)) / np.sqrt(fan_in)).astype(np.float32)
            if c.bn:
                beta = (0.1 * rng.standard_normal(c.cout)).astype(np.float32)
                gamma = rng.uniform(0.75, 1.25, c.cout).astype(np.float32)
                mean = (0.1 * rng.standard_normal(c.cout)).astype(np.float32)
                var = rng.uniform(0.75, 1.25, c.cout).astype(np.float32)
            else:
                bias = (0.1 * rng.standard_normal(c.cout)).astype(np.float32)
        else:
            raise ValueError(f"unknown init {init!r}")
        if c.bn:
            parts += [beta, gamma, mean, var, kern.ravel()]
        else:
            parts += [bias, kern.ravel()]
    return np.concatenate(parts).astype(np.float32, copy=False)


def darknet_file_bytes(stream: np.ndarray, major: int = 0, minor: int = 2, revision: int = 0, seen: int = 0) -> bytes:
    """Wrap a stream in the ``yolov3.weights`` header (yolov3_detect.py:68-79)."""
    import struct
    hdr = struct.pack("iii", major, minor, revision)
    hdr += struct.pack("q" if (major * 10 + minor) >= 2 and major < 1000 and minor < 1000 else "i", seen)
    return hdr + stream.astype("<f4").tobytes()


def images(batch: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """NHWC float32 in [0,1) — SURVEY 8d config 1 uses ``default_rng(0).random((1,416,416,3))``."""
    return np.random.default_rng(seed).random((batch, h, w, 3), dtype=np.float32)


def head_logits(batch: int, net_h: int, net_w: int, nb_class: int = 1, seed: int = 0,
                crowd: bool = False, obj_bias: float = 0.0, clamp: float = 8.0):
    """Synthetic head outputs [(B,gh,gw,C)]*3 for decode/NMS tests.

    ``crowd=True`` is BASELINE config 5: tx,ty~U(-2,2), tw,th~N(0,.5), objectness biased so
    (nearly) every candidate passes; logits clamped so exp() cannot overflow int().
    Class/objectness logits are re-drawn until the fp32 sigmoids are pairwise distinct per image
    (tie order is undefined in the reference).
    """
    rng = np.random.default_rng(seed)
    C = 3 * (5 + nb_class)
    outs = []
    for lvl in (32, 16, 8):
        gh, gw = net_h // lvl, net_w // lvl
        t = rng.standard_normal((batch, gh, gw, 3, 5 + nb_class)).astype(np.float32)
        if crowd:
            t[..., 0:2] = rng.uniform(-2, 2, t[..., 0:2].shape)
            t[..., 2:4] = 0.5 * rng.standard_normal(t[..., 2:4].shape)
        t[..., 4] += obj_bias
        np.clip(t, -clamp, clamp, out=t)
        outs.append(t.reshape(batch, gh, gw, C))
    _dedup_scores(outs, nb_class, rng)
    return outs


def _sig32(x):
    x = np.asarray(x, np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


def _dedup_scores(outs, nb_class, rng):
    B = outs[0].shape[0]
    for b in range(B):
        for c in range(nb_class):
            for _ in range(64):
                views = [o[b].reshape(-1, 5 + nb_class) for o in outs]
                s = np.concatenate([_sig32(v[:, 5 + c]) for v in views])
                _, first, cnt = np.unique(s, return_index=True, return_counts=True)
                if (cnt == 1).all():
                    break
                dup = np.ones(s.shape[0], bool); dup[first] = False
                off = 0
                for v in views:
                    m = dup[off:off + v.shape[0]]
                    v[m, 5 + c] = np.clip(rng.standard_normal(int(m.sum())), -8, 8).astype(np.float32)
                    off += v.shape[0]
            else:
                raise RuntimeError("could not de-duplicate scores")
