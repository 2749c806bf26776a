"""The letterbox oracle (oracle/letterbox.py) is pinned against the reference: the images its FaceDetector.test() loop
(src/space/face_detection.py:798-835) handed to detect(), stored by tools/make_golden.py, and cv2.resize itself."""
import os

import numpy as np
import pytest

from oracle import letterbox as LB

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "letterbox.npz")


def test_letterbox_oracle_equals_reference_loop_output():
    z = np.load(GOLDEN)
    n = len([k for k in z.files if k.startswith("src")])
    assert n == 5
    for k in range(n):
        got = LB.letterbox(z[f"src{k}"], 64)
        assert got.dtype == np.float64 and np.array_equal(got, z[f"dst{k}"]), f"case {k}"


def test_letterbox_geometry_rules():
    # face_detection.py:664-688: the longer side fills the square, the shorter one is int(short / long * S), odd padding goes below / right
    assert LB.geometry(640, 480, 416) == (416, 312, 52, 52, 0, 0)
    assert LB.geometry(375, 500, 416) == (312, 416, 0, 0, 52, 52)
    assert LB.geometry(1024, 300, 416) == (416, 121, 147, 148, 0, 0)
    assert LB.geometry(512, 512, 416) == (416, 416, 0, 0, 0, 0)


@pytest.mark.parametrize("w,h,size", [(640, 480, 416), (333, 500, 416), (97, 131, 608), (1000, 40, 416)])
def test_resize_restatement_equals_cv2(w, h, size):
    cv = pytest.importorskip("cv2")
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    w_p, h_p, pad_t, pad_b, pad_l, pad_r = LB.geometry(w, h, size)
    ref = cv.resize(img / 255, (w_p, h_p), interpolation=cv.INTER_CUBIC)
    ref = cv.copyMakeBorder(ref, pad_t, pad_b, pad_l, pad_r, cv.BORDER_CONSTANT, value=[0, 0, 0])
    assert np.array_equal(LB.letterbox(img, size), ref)
