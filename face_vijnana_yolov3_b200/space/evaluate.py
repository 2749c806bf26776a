"""Drop-in for the face-detection scorer of the reference's ``src/space/evaluate.py``: ``cal_mAP_fd`` (:27-127) and the
``cal_map_fd`` mode of its ``main`` (:337-356), SURVEY 8 row f-4.

    cal_mAP_fd(gt_path, sol_path, iou_th) -> (ps, rs, mAP)

The O(faces x detections) part - bbox_iou of every pair and the greedy "largest IoU first" assignment per image (:47-96) - runs
on the GPU in one launch over all images (``fvy_map_match``, one block per image, float64 arithmetic exactly as the CSV's numbers
reach ``bbox_iou``); reading the two CSV files, the confidence sort, the cumulative precision / recall and scipy's
``interp1d`` + ``quad`` (the reference's own library calls, :121-122) stay on the host.  Quirks kept: an image whose detections
overlap no face contributes NO rows (its false positives are not counted, :76), detections of files absent from the ground truth
are ignored, and an unmatched first image followed by matched ones raises UnboundLocalError as the reference does (:98-101).
Ties (equal IoU within an image, equal confidence) are undefined in the reference (unstable sorts); here: first in file order.
Out of scope: the identification scorers (cal_face_pairs_dists, cal_VAL_FAR, cal_acc_fi).
"""
from __future__ import annotations

import argparse

import numpy as np

from .yolov3_detect import BoundBox, bbox_iou, _box_engine   # noqa: F401  (BoundBox / bbox_iou re-exported like the reference, :17)

DEBUG = True
MODE_CAL_MAP_FD = 'cal_map_fd'


def _read(gt_path, sol_path):
    import pandas as pd
    sol_df = pd.read_csv(sol_path, header=None)           # :28
    gt_df = pd.read_csv(gt_path)                          # :34
    return gt_df, sol_df


def cal_mAP_fd(gt_path, sol_path, iou_th):
    from scipy.integrate import quad
    from scipy.interpolate import interp1d
    gt_df, sol_df = _read(gt_path, sol_path)
    gt_files = gt_df['FILE'].to_numpy()
    gt_xywh = gt_df.iloc[:, 3:7].to_numpy(np.float64)     # FACE_X, FACE_Y, FACE_WIDTH, FACE_HEIGHT (:52-56)
    sol_files = sol_df[0].to_numpy()
    sol_xywh = sol_df.iloc[:, 1:5].to_numpy(np.float64)   # x, y, w, h (:62-66)
    score = sol_df[5].to_numpy(np.float64)
    gt_rows, sol_rows = {}, {}
    for k, f in enumerate(gt_files):
        gt_rows.setdefault(f, []).append(k)
    for k, f in enumerate(sol_files):
        sol_rows.setdefault(f, []).append(k)
    images = [(k, f) for k, f in enumerate(sorted(gt_rows)) if f in sol_rows]      # groupby key order (:39); KeyError -> continue (:45-46)
    gi = np.array([r for _, f in images for r in gt_rows[f]], np.int64)
    di = np.array([r for _, f in images for r in sol_rows[f]], np.int64)
    gt_off = np.cumsum([0] + [len(gt_rows[f]) for _, f in images]).astype(np.int32)
    det_off = np.cumsum([0] + [len(sol_rows[f]) for _, f in images]).astype(np.int32)

    def boxes(xywh):                                      # BoundBox(x, y, x + w, y + h) (:52-56, :62-66)
        return np.stack([xywh[:, 0], xywh[:, 1], xywh[:, 0] + xywh[:, 2], xywh[:, 1] + xywh[:, 3]], 1) if len(xywh) else np.zeros((0, 4))
    n_boxes = max(int(len(gi) + len(di)), 2)
    det_iou, img_any = _box_engine(n_boxes, 1).map_match(boxes(gt_xywh[gi]), gt_off, boxes(sol_xywh[di]), det_off)
    res_score, res_iou, started = [], [], False
    for n, (k, f) in enumerate(images):
        if not img_any[n]:
            continue                                      # :76
        if k != 0 and not started:
            raise UnboundLocalError("cannot access local variable 'res_df' where it is not associated with a value")   # :98-101
        started = True
        lo, hi = det_off[n], det_off[n + 1]
        res_score.append(score[di[lo:hi]]); res_iou.append(det_iou[lo:hi])
    if not started:
        raise UnboundLocalError("cannot access local variable 'res_df' where it is not associated with a value")
    res_score, res_iou = np.concatenate(res_score), np.concatenate(res_iou)
    order = np.argsort(-res_score, kind='stable')         # confidence descending (:105)
    tp = np.cumsum(res_iou[order] >= iou_th)              # :114-118
    ps = tp / np.arange(1, len(order) + 1)
    rs = tp / gt_df.shape[0]
    func = interp1d(rs, ps)                               # :121
    mAP = quad(lambda x: func(x), rs[0], rs[-1])          # :122
    return ps, rs, mAP[0]


def main(args):
    """The cal_map_fd mode of the reference's main (:337-356); the curves are written as p_r_curve.h5 with h5lite."""
    if args.mode != MODE_CAL_MAP_FD:
        raise SystemExit(f"mode {args.mode!r} is outside the face-detection hot path (only {MODE_CAL_MAP_FD!r} is provided)")
    ps_ls, rs_ls, mAP_ls = [], [], []
    for iou_th in np.arange(0.5, 1.0, 0.05):
        ps, rs, mAP = cal_mAP_fd(args.gt_path, args.sol_path, iou_th)
        if DEBUG:
            print('{0:1.2f}'.format(iou_th), mAP)
        ps_ls.append(ps); rs_ls.append(rs); mAP_ls.append(mAP)
    from .. import h5lite
    h5lite.write_h5('p_r_curve.h5', {'ps_ls': np.asarray(ps_ls), 'rs_ls': np.asarray(rs_ls), 'mAP_ls': np.asarray(mAP_ls)})


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--mode')
    parser.add_argument('--gt_path')
    parser.add_argument('--sol_path')
    main(parser.parse_args())
