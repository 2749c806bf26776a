"""ORACLE - test infrastructure only.  CPU restatement of the reference's face-detection scorer ``cal_mAP_fd``
(/root/reference/src/space/evaluate.py:27-127).  Pinned: tests/golden/map_fd.npz holds the inputs and the outputs of the REAL
reference function (executed by tools/make_golden.py::map_fd_cases under the one-line pandas shim of oracle/ref_loader.py), and
tests/test_oracle_pin.py runs the live reference beside this file when /root/reference is present.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import csv
import math
from typing import Dict, List, Tuple

import numpy as np


def _interval_overlap(x1, x2, x3, x4):            # yolov3_detect.py:165-178
    if x3 < x1:
        if x4 < x1:
            return 0
        return min(x2, x4) - x1
    if x2 < x3:
        return 0
    return min(x2, x4) - x3


def bbox_iou(a, b) -> float:                       # yolov3_detect.py:183-194 on (x1, y1, x2, y2); zero union -> nan like NumPy
    iw = _interval_overlap(a[0], a[2], b[0], b[2])
    ih = _interval_overlap(a[1], a[3], b[1], b[3])
    inter = iw * ih
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    if union == 0:
        return math.nan if inter == 0 else math.copysign(math.inf, inter)
    return float(inter) / union


def read_tables(gt_path: str, sol_path: str):
    """gt: header FACE_ID, FILE, SUBJECT_ID, FACE_X, FACE_Y, FACE_WIDTH, FACE_HEIGHT (evaluate.py:34, :52-56 use columns 1, 3..6);
    sol: no header, file, x, y, w, h, score (:28, :62-66).  Numbers are parsed as pandas would: int64 if every value of the column is
    an integer literal, else float64."""
    def col(values: List[str]):
        try:
            return np.array([int(v) for v in values], np.int64).astype(np.float64)
        except ValueError:
            return np.array([float(v) for v in values], np.float64)
    with open(gt_path, newline="") as f:
        rows = list(csv.reader(f))[1:]
    gt_files = [r[1] for r in rows]
    gt = np.stack([col([r[c] for r in rows]) for c in (3, 4, 5, 6)], 1) if rows else np.zeros((0, 4))
    with open(sol_path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    sol_files = [r[0] for r in rows]
    sol = np.stack([col([r[c] for r in rows]) for c in (1, 2, 3, 4, 5)], 1) if rows else np.zeros((0, 5))
    return gt_files, gt, sol_files, sol


def match_image(gt_xywh: np.ndarray, det_xywh: np.ndarray) -> Tuple[np.ndarray, bool]:
    """evaluate.py:47-96 for one image -> (IoU assigned to each detection or -1, any pair with IoU > 0)."""
    out = np.full(len(det_xywh), -1.0)
    pairs = []
    for i, g in enumerate(gt_xywh):
        gb = (g[0], g[1], g[0] + g[2], g[1] + g[3])
        for j, d in enumerate(det_xywh):
            v = bbox_iou(gb, (d[0], d[1], d[0] + d[2], d[1] + d[3]))
            if v > 0.0:
                pairs.append((i, j, v))
    if not pairs:
        return out, False
    pairs.sort(key=lambda t: (-t[2], t[0], t[1]))                      # IoU descending; ties: (i, j) ascending (framework's rule)
    while pairs:
        i, j, v = pairs[0]
        out[j] = v
        pairs = [p for p in pairs if p[0] != i and p[1] != j]
    return out, True


def cal_mAP_fd(gt_path: str, sol_path: str, iou_th: float):
    """-> (ps, rs, mAP) exactly as the reference returns them."""
    from scipy.integrate import quad
    from scipy.interpolate import interp1d
    gt_files, gt, sol_files, sol = read_tables(gt_path, sol_path)
    by_gt: Dict[str, List[int]] = {}
    for k, f in enumerate(gt_files):
        by_gt.setdefault(f, []).append(k)
    by_sol: Dict[str, List[int]] = {}
    for k, f in enumerate(sol_files):
        by_sol.setdefault(f, []).append(k)
    res_rows: List[Tuple[float, float]] = []             # (score, assigned IoU) in the reference's concatenation order
    started = False
    for k, image_id in enumerate(sorted(by_gt)):          # groupby sorts its keys (:39)
        if image_id not in by_sol:
            continue                                      # :45-46
        det = sol[by_sol[image_id]]
        ious, any_pair = match_image(gt[by_gt[image_id]], det[:, :4])
        if not any_pair:
            continue                                      # :76: the image's detections never reach res_df
        if k != 0 and not started:
            raise UnboundLocalError("cannot access local variable 'res_df' where it is not associated with a value")   # :98-101 quirk
        started = True
        res_rows += [(det[j, 4], ious[j]) for j in range(len(det))]
    if not started:
        raise UnboundLocalError("cannot access local variable 'res_df' where it is not associated with a value")
    order = sorted(range(len(res_rows)), key=lambda r: -res_rows[r][0])     # confidence descending (:105); ties keep concat order
    ps, rs, tp = [], [], 0
    for count, r in enumerate(order, 1):
        if res_rows[r][1] >= iou_th:
            tp += 1
        ps.append(tp / count)
        rs.append(tp / len(gt_files))
    ps, rs = np.asarray(ps), np.asarray(rs)
    func = interp1d(rs, ps)
    mAP = quad(lambda x: func(x), rs[0], rs[-1])
    return ps, rs, mAP[0]
