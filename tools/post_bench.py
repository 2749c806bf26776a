"""BASELINE.json configs[4] - NMS / decode stress: dense synthetic crowd images, all nine anchors, ~10 647 candidates per image
(416x416), nms_iou_th = 0.5, num_cands = 60, batch 40.  Head logits resident in HBM; one step = decode_netout +
correct_yolo_boxes + do_nms + selection for the batch (fvy_postprocess).  Prints one JSON line: images/s, the step time from
CUDA events on the handle's stream, algorithmic bytes per SURVEY 8d (decode: logits read + records written; NMS: sorted
records + the n^2/8-byte suppression bitmask written and read) against the measured HBM peak, and the C oracle timed on one
image on the host.  Per-kernel times / DRAM bytes come from an ncu pass over the same command (tools/evidence_post.sh).

usage: python tools/post_bench.py [--batch 40] [--steps 20] [--warmup 5]
"""
import argparse, json, os, sys, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from face_vijnana_yolov3_b200 import _lib as L, synth          # noqa: E402
from face_vijnana_yolov3_b200.engine import Engine, post_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=40)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    import torch
    B, S = a.batch, a.size
    outs = synth.head_logits(B, S, S, 1, seed=4, crowd=True, obj_bias=6.0)
    dev = [torch.from_numpy(np.ascontiguousarray(o)).cuda() for o in outs]
    eng = Engine(S, S, head=L.HEAD_NONE, nb_class=1, max_batch=B)
    pp = post_params(0.5, 0.5, anchor_mask=L.ANCHOR_MASK_ALL, num_cands=60)
    hw = np.array([[S, S]] * B, np.int32)
    for _ in range(a.warmup):
        dets, counts = eng.postprocess(dev, pp=pp, image_hw=hw, max_out=60)
    l0 = eng.launch_count
    eng.timer_start()
    for _ in range(a.steps):
        dets, counts = eng.postprocess(dev, pp=pp, image_hw=hw, max_out=60)
    ms = eng.timer_stop() / a.steps
    launches = (eng.launch_count - l0) / a.steps
    # candidates per image: every cell x anchor passes with obj_bias 6 (checked by tests/test_gpu_parity.py::test_nms_stress_crowd_10k)
    n = sum(3 * (S // s) * (S // s) for s in (32, 16, 8))
    logits_bytes = n * 6 * 4
    decode_bytes = logits_bytes + 28 * n
    words = (n + 63) // 64
    nms_bytes = 24 * n + 2 * n * words * 8
    alg = B * (decode_bytes + nms_bytes)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    achieved = alg / (ms * 1e-3) / 1e9
    line = {"metric": "images/sec (decode+NMS stress, ~10k candidates/image)", "value": B / (ms * 1e-3), "unit": "images/s", "n_gpus": 1,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "int64/f64 IoU, f32 scores", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: crowd logits, all 9 anchors, {n} candidates/image, batch {B} @{S}, nms_iou_th 0.5, num_cands 60",
                       "pairs_per_image": n * (n - 1) // 2},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "algorithmic_bytes_per_step": alg, "decode_bytes_per_image": decode_bytes, "nms_bytes_per_image": nms_bytes,
                         "note": "the step is five kernels; the bitmask kernel (n^2/2 IoU tests per image) is compute-bound, the others latency-bound "
                                 "per image - see the ncu per-kernel list"},
            "gpu_launches_per_step": launches, "kept_per_image_mean": float(np.mean(counts))}
    if not a.no_cpu_baseline:
        from oracle import postproc as P
        t0 = time.perf_counter()
        dd = P.decode_image([o[0] for o in outs], anchor_masks=P.ALL_ANCHOR_MASK, obj_thresh=0.5, net_h=S, net_w=S)
        ib = P.correct_yolo_boxes(dd["box"], S, S, S, S)
        P.do_nms(ib, dd["classes"], 0.5)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "images/s", "cores": 1, "kind": "port", "candidates": int(len(ib)),
                                "sample": f"1 image of the same workload through the C oracle (decode + correct + do_nms), {dt:.2f} s"}
    print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
