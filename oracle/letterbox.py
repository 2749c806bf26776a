"""ORACLE (test infrastructure, not product): CPU restatement of the pre-processing of FaceDetector.evaluate / test,
/root/reference/src/space/face_detection.py:657-690 and :798-835

    image = imread(file) / 255
    image = cv.resize(image, (w_p, h_p), interpolation=cv.INTER_CUBIC)
    image = cv.copyMakeBorder(image, pad_t, pad_b, pad_l, pad_r, cv.BORDER_CONSTANT, value=[0, 0, 0])

`cv.resize` lives in a third-party dependency that is not vendored under /root/reference (setup.py:19 pins
opencv-contrib-python==4.2.0.32; this image has opencv 4.13.0).  Its published algorithm for CV_64F / INTER_CUBIC
(modules/imgproc/src/resize.cpp: resizeGeneric_<HResizeCubic<double,double,float>, VResizeCubic<double,double,float,...>>,
interpolateCubic with A = -0.75) is restated here in numpy:
  * source coordinate of destination index d:  f = float32((d + 0.5) * scale - 0.5), scale = 1 / (dst / src) in float64;
    s = floor(f); x = f - s in float32;
  * taps s-1 .. s+2 clamped to [0, src-1] (replicated border);
  * coefficients in float32, evaluated exactly as interpolateCubic writes them;
  * horizontal pass first, then vertical, both accumulating in float64 left to right.
PINNED: tests/test_oracle_letterbox.py checks this restatement bit for bit against cv2.resize itself (present in this
image and on the GPU box) on seeded images, and against tests/golden/letterbox.npz generated from the reference's own lines by
tools/make_golden.py.
"""
from __future__ import annotations

import numpy as np


def geometry(w: int, h: int, size: int):
    """(w_p, h_p, pad_t, pad_b, pad_l, pad_r) as face_detection.py:664-688 computes them (Python float arithmetic)."""
    pad_t = pad_b = pad_l = pad_r = 0
    if w >= h:
        w_p = size
        h_p = int(h / w * size)
        pad = size - h_p
        pad_t, pad_b = pad // 2, pad // 2 + (pad % 2)
    else:
        h_p = size
        w_p = int(w / h * size)
        pad = size - w_p
        pad_l, pad_r = pad // 2, pad // 2 + (pad % 2)
    return w_p, h_p, pad_t, pad_b, pad_l, pad_r


def _cubic_coeffs(x: np.ndarray) -> np.ndarray:
    f = np.float32
    A, x, one = f(-0.75), x.astype(np.float32), f(1)
    c0 = ((A * (x + one) - f(5) * A) * (x + one) + f(8) * A) * (x + one) - f(4) * A
    c1 = ((A + f(2)) * x - (A + f(3))) * x * x + one
    c2 = ((A + f(2)) * (one - x) - (A + f(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def _axis_table(dst: int, src: int):
    scale = 1.0 / (dst / src)
    d = np.arange(dst)
    fpos = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(fpos).astype(np.int64)
    frac = (fpos - s.astype(np.float32)).astype(np.float32)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, src - 1)
    return idx, _cubic_coeffs(frac).astype(np.float64)


def resize_cubic(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv.resize(img, (dst_w, dst_h), interpolation=cv.INTER_CUBIC) for a float64 (H, W, C) image."""
    img = np.asarray(img, np.float64)
    xi, xa = _axis_table(dst_w, img.shape[1])
    yi, ya = _axis_table(dst_h, img.shape[0])
    rows = img[:, xi[:, 0], :] * xa[:, 0][None, :, None]
    for k in range(1, 4):
        rows = rows + img[:, xi[:, k], :] * xa[:, k][None, :, None]
    out = rows[yi[:, 0]] * ya[:, 0][:, None, None]
    for k in range(1, 4):
        out = out + rows[yi[:, k]] * ya[:, k][:, None, None]
    return out


def letterbox(image_u8: np.ndarray, size: int) -> np.ndarray:
    """The (size, size, 3) float64 image the reference hands to FaceDetector.detect for a uint8 RGB image."""
    h, w = image_u8.shape[0], image_u8.shape[1]
    w_p, h_p, pad_t, pad_b, pad_l, pad_r = geometry(w, h, size)
    out = np.zeros((size, size, 3), np.float64)
    out[pad_t:pad_t + h_p, pad_l:pad_l + w_p] = resize_cubic(image_u8 / 255, w_p, h_p)
    return out
