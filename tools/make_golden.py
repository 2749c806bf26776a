"""Generate tests/golden/* by running the REAL reference in the build container.

/root/reference does not exist on the GPU box, so the outputs of its own functions are committed
as small fixtures (this script is their provenance):

  darknet53_summary.json  conv/bnorm/add rows of the Keras summary() dump the reference ships in
                          analysis/face_recog_analysis.ipynb:1431-1868 (+ the param totals).
  post_yolo3_*.npz        reference decode_netout -> correct_yolo_boxes -> do_nms on seeded logits
                          (src/space/yolov3_detect.py:335-444), executed with this container's NumPy
                          (>= 2, i.e. float32 scalar arithmetic = FVY_ARITH_F32).
  post_fd6.npz            reference FaceDetector.detect (src/space/face_detection.py:885-949) on seeded
                          (1,13,13,6) maps through a fake model.predict.
  iou_cases.json          bbox_iou / _interval_overlap known answers computed by the reference code.
  post_yolo3_mc.npz       the same pipeline with nb_class = 3: pins the per-class loop of do_nms (:431-444) on the reference itself.
  post_float.npz          reference bbox_iou / do_nms called on FLOAT boxes (before correct_yolo_boxes): np.float32 coordinates as this
                          NumPy's decode_netout returns them (float32 arithmetic) and the same boxes as Python floats (float64).
  map_fd.npz              reference evaluate.cal_mAP_fd (src/space/evaluate.py:27-127) on synthetic ground-truth / detection CSVs: the
                          CSV texts and the (ps, rs, mAP) it returns for several IoU thresholds (run under the one-line pandas shim
                          of oracle/ref_loader.py::load_evaluate).
  letterbox.npz           the image the reference's own FaceDetector.test() loop (src/space/face_detection.py:798-835, cv2 of this
                          image) hands to detect(), for seeded uint8 images written as PNG-exact BMP files, image_size 64
  gt_tensor.npz           reference TrainingSequence.__getitem__ ground-truth tensors (src/space/face_detection.py:98-310)
                          for synthetic images / training.csv rows (letterbox geometry + cell assignment).

usage: python tools/make_golden.py
"""
from __future__ import annotations

import json
import os
import re
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from face_vijnana_yolov3_b200 import synth   # noqa: E402
from oracle import ref_loader as R            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ANCHORS = [[116, 90, 156, 198, 373, 326], [30, 61, 62, 45, 59, 119], [10, 13, 16, 30, 33, 23]]


def summary_fixture():
    path = "/root/reference/analysis/face_recog_analysis.ipynb"
    lines = open(path).read().split("\n")[1425:1875]
    rows = []
    pat = re.compile(r'"(\w+) \((\w+)\)\s+\(None, (\d+), (\d+), (\d+)\)?\s+(\d+)')
    for ln in lines:
        m = pat.search(ln)
        if m and m.group(2) in ("Conv2D", "BatchNormalization", "Add"):
            rows.append([m.group(1), m.group(2), int(m.group(3)), int(m.group(4)), int(m.group(5)), int(m.group(6))])
    tot = [int(re.search(r"([\d,]+)", ln.split(":")[1]).group(1).replace(",", "")) for ln in lines if "params:" in ln][:3]
    json.dump({"source": "analysis/face_recog_analysis.ipynb:1431-1868", "rows": rows, "total": tot[0], "trainable": tot[1],
               "non_trainable": tot[2]}, open(os.path.join(OUT, "darknet53_summary.json"), "w"))
    print("summary rows", len(rows), tot)


def post_yolo3(tag, seed, image_hw, obj_thresh, nms_thresh, nb_class=1, all_anchors=False):
    Y = R.load_yolov3_detect()
    outs = synth.head_logits(1, 416, 416, nb_class, seed=seed)
    boxes = []
    src = Y.decode_netout
    for i in range(3):
        boxes += src(outs[i][0].copy(), ANCHORS[i], i, obj_thresh, 416, 416)
    nbox = np.array([[b.xmin, b.ymin, b.xmax, b.ymax] for b in boxes], np.float64)
    objn = np.array([b.objness for b in boxes], np.float32)
    cls0 = np.array([np.array(b.classes, np.float32) for b in boxes], np.float32).reshape(len(boxes), nb_class)
    Y.correct_yolo_boxes(boxes, image_hw[0], image_hw[1], 416, 416)
    ibox = np.array([[b.xmin, b.ymin, b.xmax, b.ymax] for b in boxes], np.int64)
    Y.do_nms(boxes, nms_thresh)
    cls1 = np.array([np.array(b.classes, np.float32) for b in boxes], np.float32).reshape(len(boxes), nb_class)
    np.savez_compressed(os.path.join(OUT, f"post_yolo3_{tag}.npz"), out0=outs[0][0], out1=outs[1][0], out2=outs[2][0],
                        nbox=nbox, objness=objn, classes_before=cls0, ibox=ibox, classes_after=cls1,
                        image_hw=np.array(image_hw, np.int32), obj_thresh=obj_thresh, nms_thresh=nms_thresh,
                        numpy_version=np.__version__)
    print(tag, "cands", len(boxes), "kept", int((cls1 > 0).any(1).sum()))


def post_fd6():
    rng = np.random.default_rng(11)
    maps, res = [], []
    hps = {"face_conf_th": 0.5, "nms_iou_th": 0.5, "num_cands": 60}
    for k in range(4):
        m = rng.standard_normal((1, 13, 13, 6)).astype(np.float32)
        m[..., 0] += 1.5; m[..., 5] += 1.5
        m[..., 1:3] = rng.uniform(-0.2, 1.2, m[..., 1:3].shape)
        m[..., 3:5] = rng.uniform(-0.05, 0.5, m[..., 3:5].shape) * (1 + 2 * (k % 2))
        if k == 3:
            hps = {"face_conf_th": 0.3, "nms_iou_th": 0.3, "num_cands": 10}
        fd = R.make_ref_face_detector(hps, 416, predict_fn=lambda image, m=m: m.copy())
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            boxes = fd.detect(np.zeros((1, 416, 416, 3)))
        r = np.array([[b.xmin, b.ymin, b.xmax, b.ymax] for b in boxes], np.int64).reshape(-1, 4)
        sc = np.array([b.classes[0] for b in boxes], np.float32)
        ob = np.array([b.objness for b in boxes], np.float32)
        maps.append(m[0]); res.append((r, sc, ob, dict(hps)))
        print("fd6 case", k, "returned", len(boxes))
    np.savez_compressed(os.path.join(OUT, "post_fd6.npz"), maps=np.stack(maps),
                        **{f"ibox{k}": res[k][0] for k in range(4)}, **{f"score{k}": res[k][1] for k in range(4)},
                        **{f"obj{k}": res[k][2] for k in range(4)},
                        hps=json.dumps([res[k][3] for k in range(4)]))


def iou_cases():
    Y = R.load_yolov3_detect()
    rng = np.random.default_rng(5)
    cases = [[0, 0, 10, 10, 5, 5, 15, 15], [0, 0, 10, 10, 10, 10, 20, 20], [0, 0, 10, 10, 0, 0, 10, 10], [0, 0, 10, 10, 2, 2, 4, 4],
             [-5, -5, 5, 5, 0, 0, 3, 30], [0, 0, 1, 1, 5, 5, 6, 6], [3, 3, 3, 9, 3, 3, 8, 9], [100, 50, 300, 90, 90, 60, 110, 70]]
    for _ in range(56):
        a = rng.integers(-50, 400, 4); b = rng.integers(-50, 400, 4)
        cases.append([int(min(a[0], a[2])), int(min(a[1], a[3])), int(max(a[0], a[2])), int(max(a[1], a[3])),
                      int(min(b[0], b[2])), int(min(b[1], b[3])), int(max(b[0], b[2])), int(max(b[1], b[3]))])
    out = []
    for c in cases:
        b1, b2 = Y.BoundBox(*c[:4]), Y.BoundBox(*c[4:])
        try:
            v = Y.bbox_iou(b1, b2)
        except ZeroDivisionError:
            v = None
        out.append({"a": c[:4], "b": c[4:], "iou": v,
                    "ow": int(Y._interval_overlap([c[0], c[2]], [c[4], c[6]])), "oh": int(Y._interval_overlap([c[1], c[3]], [c[5], c[7]]))})
    json.dump(out, open(os.path.join(OUT, "iou_cases.json"), "w"))
    print("iou cases", len(out))


def float_box_cases():
    """bbox_iou / do_nms are type-generic (:165-194, :426-444): run them on un-corrected float boxes in both arithmetic types."""
    Y = R.load_yolov3_detect()
    outs = synth.head_logits(1, 416, 416, 1, seed=31)
    boxes = []
    for i in range(3):
        boxes += Y.decode_netout(outs[i][0].copy(), ANCHORS[i], i, 0.6, 416, 416)
    assert all(isinstance(b.xmin, np.float32) for b in boxes), "this NumPy returns float32 coordinates"
    box32 = np.array([[b.xmin, b.ymin, b.xmax, b.ymax] for b in boxes], np.float32)
    cls0 = np.array([b.classes[0] for b in boxes], np.float32)
    rng = np.random.default_rng(3)
    pairs = rng.integers(0, len(boxes), (96, 2))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        iou32 = np.array([Y.bbox_iou(boxes[i], boxes[j]) for i, j in pairs], np.float64)
        assert all(isinstance(Y.bbox_iou(boxes[i], boxes[j]), np.float32) for i, j in pairs[:4])
        pyb = [Y.BoundBox(float(b.xmin), float(b.ymin), float(b.xmax), float(b.ymax), b.objness, np.array(b.classes, np.float32)) for b in boxes]
        iou64 = []
        for i, j in pairs:
            try:
                iou64.append(Y.bbox_iou(pyb[i], pyb[j]))
            except ZeroDivisionError:
                iou64.append(np.nan)
        b32 = [Y.BoundBox(b.xmin, b.ymin, b.xmax, b.ymax, b.objness, np.array(b.classes, np.float32)) for b in boxes]
        Y.do_nms(b32, 0.45)
        Y.do_nms(pyb, 0.45)
    after32 = np.array([b.classes[0] for b in b32], np.float32)
    after64 = np.array([b.classes[0] for b in pyb], np.float32)
    np.savez_compressed(os.path.join(OUT, "post_float.npz"), box=box32, classes_before=cls0, pairs=pairs, iou32=iou32,
                        iou64=np.array(iou64, np.float64), after32=after32, after64=after64, nms_thresh=0.45, numpy_version=np.__version__)
    print("float boxes", len(boxes), "kept f32", int((after32 > 0).sum()), "kept f64", int((after64 > 0).sum()),
          "differ", int(((after32 > 0) != (after64 > 0)).sum()))


def map_fd_cases():
    """Reference cal_mAP_fd on two synthetic result sets: float detections (what FaceDetector.test writes after the un-letterbox
    scaling) and integer-valued ones (pandas parses int64: the integer IoU path).  Covers: several faces / detections per image,
    an image without detections, an image whose detections overlap no face (skipped by the reference, evaluate.py:76), a detection
    file absent from the ground truth."""
    import tempfile
    import pandas as pd
    E = R.load_evaluate()
    out = {}
    for tag, as_int in (("f", False), ("i", True)):
        rng = np.random.default_rng(3 if not as_int else 4)
        rows, det, fid = [], [], 0
        for k in range(12):
            f = f"im{k:02d}.jpg"
            gts = []
            for _ in range(int(rng.integers(1, 6))):
                x, y = rng.integers(0, 300, 2); w, h = rng.integers(20, 120, 2)
                rows.append([fid, f, 1000 + fid, int(x), int(y), int(w), int(h)]); fid += 1; gts.append((x, y, w, h))
            if k == 4:
                continue                                            # no detections for this image
            for d in range(int(rng.integers(1, 9))):
                if k == 7:
                    x, y, w, h = 1000 + 10 * d, 1000, 20, 20        # overlaps nothing: the image is skipped
                elif d < len(gts) and rng.random() < 0.8:
                    x, y, w, h = gts[d]; x = x + rng.normal(0, 6); y = y + rng.normal(0, 6); w = w * rng.uniform(0.8, 1.2); h = h * rng.uniform(0.8, 1.2)
                else:
                    x, y = rng.uniform(0, 300, 2); w, h = rng.uniform(20, 120, 2)
                v = [int(round(x)), int(round(y)), int(round(w)), int(round(h))] if as_int else [float(x), float(y), float(w), float(h)]
                det.append([f] + v + [float(rng.random())])
        det.append(["not_in_gt.jpg", 5, 5, 50, 50, 0.9] if as_int else ["not_in_gt.jpg", 5.0, 5.0, 50.0, 50.0, 0.9])
        with tempfile.TemporaryDirectory() as d:
            gt_path, sol_path = os.path.join(d, "gt.csv"), os.path.join(d, "sol.csv")
            pd.DataFrame(rows, columns=["FACE_ID", "FILE", "SUBJECT_ID", "FACE_X", "FACE_Y", "FACE_WIDTH", "FACE_HEIGHT"]).to_csv(gt_path, index=False)
            pd.DataFrame(det).to_csv(sol_path, index=False, header=False)
            out[f"gt_csv_{tag}"] = open(gt_path).read(); out[f"sol_csv_{tag}"] = open(sol_path).read()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for th in (0.5, 0.75):
                    ps, rs, mAP = E.cal_mAP_fd(gt_path, sol_path, th)
                    out[f"ps_{tag}_{int(th * 100)}"] = ps; out[f"rs_{tag}_{int(th * 100)}"] = rs; out[f"mAP_{tag}_{int(th * 100)}"] = mAP
                    print("map_fd", tag, th, "rows", len(ps), "mAP", mAP)
    np.savez_compressed(os.path.join(OUT, "map_fd.npz"), **out)


def gt_tensor_cases():
    """Reference TrainingSequence.__getitem__ (src/space/face_detection.py:98-310) on synthetic images + training.csv."""
    import tempfile
    import cv2 as cv
    import pandas as pd
    fdm = R.load_face_detection()
    fdm.imread = lambda path: cv.imread(path, cv.IMREAD_COLOR)[:, :, ::-1]
    sizes = {"a_wide.jpg": (640, 480), "b_tall.jpg": (375, 500), "c_square.jpg": (512, 512), "d_wide2.jpg": (1024, 300), "e_tall2.jpg": (200, 711)}
    rng = np.random.default_rng(5)
    rows = []
    fid = 0
    for name, (w, h) in sizes.items():
        for k in range(6):
            fw, fh = int(rng.integers(8, max(9, w // 3))), int(rng.integers(8, max(9, h // 3)))
            fx, fy = int(rng.integers(1, w - fw)), int(rng.integers(1, h - fh))
            if k == 4:
                fw = 0                                      # invalid row: skipped by the reference (:147-149)
            rows.append([fid, name, 1000 + fid, fx, fy, fw, fh]); fid += 1
    with tempfile.TemporaryDirectory() as d:
        for name, (w, h) in sizes.items():
            cv.imwrite(os.path.join(d, name), np.zeros((h, w, 3), np.uint8))
        pd.DataFrame(rows, columns=["FACE_ID", "FILE", "SUBJECT_ID", "FACE_X", "FACE_Y", "FACE_WIDTH", "FACE_HEIGHT"]).to_csv(
            os.path.join(d, "training.csv"), index=False)
        seq = fdm.FaceDetector.TrainingSequence(d, {"batch_size": 2}, {"image_size": 416, "bb_info_c_size": 6}, 13, 32)
        names, targets, shapes = list(seq.file_names), [], []
        for i in range(len(seq)):
            x, y = seq[i]
            targets += list(y["output"]); shapes += [im.shape for im in x["input1"]]
    assert len(targets) == len(names) and all(s == (416, 416, 3) for s in shapes)
    faces = np.array([[r[3], r[4], r[5], r[6]] for r in rows], np.int64)
    owner = np.array([names.index(r[1]) for r in rows], np.int64)
    wh = np.array([sizes[n] for n in names], np.int64)
    np.savez_compressed(os.path.join(OUT, "gt_tensor.npz"), faces=faces, owner=owner, wh=wh, targets=np.asarray(targets, np.float64))
    print("gt tensors", len(targets), "positive cells", int(sum((t[..., 0] > 0).sum() for t in targets)))


def letterbox_cases():
    """Reference FaceDetector.test() (src/space/face_detection.py:783-883) run with a capturing detect(): its own imread / 255,
    cv.resize(INTER_CUBIC) and cv.copyMakeBorder lines produce the images stored here."""
    import tempfile
    import cv2 as cv
    fdm = R.load_face_detection()
    fdm.DEBUG = False
    fdm.imread = lambda path: cv.imread(path, cv.IMREAD_COLOR)[:, :, ::-1]
    sizes = [(131, 97), (97, 131), (80, 80), (200, 60), (33, 150)]          # (w, h): down- and up-scaling, both orientations
    rng = np.random.default_rng(11)
    captured = {}
    fd = fdm.FaceDetector.__new__(fdm.FaceDetector)
    fd.nn_arch = {"image_size": 64, "bb_info_c_size": 6}
    fd.hps = {"face_conf_th": 0.5, "nms_iou_th": 0.5, "num_cands": 60}
    srcs = []
    with tempfile.TemporaryDirectory() as d:
        fd.conf = {"test_path": d, "output_file_path": os.path.join(d, "out.csv")}
        for k, (w, h) in enumerate(sizes):
            im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            cv.imwrite(os.path.join(d, f"im{k}.bmp.jpg"), im[:, :, ::-1], [cv.IMWRITE_JPEG_QUALITY, 100])
            srcs.append(cv.imread(os.path.join(d, f"im{k}.bmp.jpg"), cv.IMREAD_COLOR)[:, :, ::-1].copy())     # the pixels the loop will see
        order = []

        def fake_detect(image):
            captured[len(order)] = np.array(image[0]); order.append(1)
            return []
        fd.detect = fake_detect
        import glob as _glob
        names = _glob.glob(os.path.join(d, "*.jpg"))
        fd.test()
        idx = [int(os.path.basename(n)[2]) for n in names]
    out = {}
    for pos, k in enumerate(idx):
        out[f"src{k}"] = srcs[k]; out[f"dst{k}"] = captured[pos]
    np.savez_compressed(os.path.join(OUT, "letterbox.npz"), **out)
    print("letterbox cases", len(idx), [captured[i].shape for i in range(len(idx))])


if __name__ == "__main__":
    if not R.available():
        raise SystemExit("/root/reference is not available: golden fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    summary_fixture()
    post_yolo3("a", seed=21, image_hw=(416, 416), obj_thresh=0.5, nms_thresh=0.45)
    post_yolo3("b", seed=22, image_hw=(360, 640), obj_thresh=0.6, nms_thresh=0.5)
    post_yolo3("c", seed=23, image_hw=(500, 375), obj_thresh=0.5, nms_thresh=0.3)
    post_yolo3("mc", seed=25, image_hw=(416, 416), obj_thresh=0.6, nms_thresh=0.45, nb_class=3)
    float_box_cases()
    map_fd_cases()
    post_fd6()
    iou_cases()
    gt_tensor_cases()
    letterbox_cases()
