#!/usr/bin/env python
"""Headline benchmark: images/sec @416 for forward + decode + NMS (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...   (the reference's CPU path, oracle port)

One step = one pass of the hot path (Darknet-53/YOLOv3 forward -> decode_netout/correct_yolo_boxes
-> do_nms) over one batch of 40 synthetic 416x416 images per GPU (BASELINE.json configs[1]).
`value` is timed with inputs resident in HBM; `e2e` goes through the public API with host buffers.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @416 (fwd+decode+NMS)"
UNIT = "images/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def mark(self):
        """Samples read so far are discarded when the summary is made (they precede the load)."""
        self.first = len(self.rows)

    def start(self):
        self.first = 0
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.rows[self.first:]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_pass(n_images, size, seed, threads):
    """The reference's CPU path restated (oracle port): torch-CPU fp32 forward of make_yolov3_model
    (yolov3_detect.py:217-311) + decode_netout/correct_yolo_boxes/do_nms (:335-444).  Returns seconds."""
    import torch
    from face_vijnana_yolov3_b200 import arch, synth
    from oracle import darknet_ref as D, postproc as P
    torch.set_num_threads(threads)
    stream = synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT)
    x = synth.images(n_images, size, size, seed)
    t0 = time.perf_counter()
    outs = D.forward(stream, x, 1)
    kept = 0
    for b in range(n_images):
        d = P.decode_image([o[b] for o in outs], obj_thresh=0.5, net_h=size, net_w=size)
        ib = P.correct_yolo_boxes(d["box"], size, size, size, size)
        cls = P.do_nms(ib, d["classes"], 0.45)
        kept += int((cls > 0).any(1).sum())
    return time.perf_counter() - t0, kept


def run_reference(args, rank):
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = args.ref_images
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_pass(1, args.size, 0, threads)
    t = 0.0
    for s in range(args.steps):
        dt, _ = cpu_reference_pass(sample, args.size, s, threads)
        t += dt
    v = sample * args.steps / t
    line = {"impl": "reference", "metric": METRIC.replace("@416", f"@{args.size}"), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"YOLOv3 face detector (nb_class=1) inference batch {args.batch} @{args.size}x{args.size}: forward+decode_netout+do_nms",
                       "net": args.size, "batch_per_gpu": args.batch, "sample_images_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} of the {args.batch} images per step x {args.steps} steps; torch-CPU fp32 restatement of the Keras graph "
                                       "+ C restatement of decode/NMS (Keras/TF not installable; reference is pure Python)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version banner to stdout when the
    process group comes up), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the real stdout."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_OUT = sys.stdout


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="fvy", choices=["fvy", "reference"])
    ap.add_argument("--batch", type=int, default=40, help="images per GPU per step (BASELINE configs[1])")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--ref-images", type=int, default=2, help="images per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tile-n", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "fvy":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from face_vijnana_yolov3_b200 import _lib as L, arch, synth
    from face_vijnana_yolov3_b200.engine import DET_DTYPE, Engine, post_params
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the CUDA hot path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("FVY_NCCL_DEBUG", "WARN")   # NCCL prints its version banner on stdout otherwise
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, S = args.batch, args.size
    eng = Engine(S, S, head=L.HEAD_YOLO3, nb_class=1, max_batch=B, device=local_rank, tile_n_max=args.tile_n)
    eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
    pp = post_params(obj_thresh=0.5, nms_thresh=0.45)
    max_out = sum(g[0] * g[1] * n for g, n in zip(eng.grids, (1, 2, 1)))   # most candidates the reference's anchor mask can produce (4225 @416)
    x_host = torch.from_numpy(synth.images(B, S, S, seed=1000 + rank)).pin_memory()
    x_dev = x_host.cuda(non_blocking=False)
    hw_dev = torch.tensor([[S, S]] * B, dtype=torch.int32, device="cuda")
    dets_dev = torch.empty((B, max_out, 8), dtype=torch.int32, device="cuda")
    cnt_dev = torch.empty((B,), dtype=torch.int32, device="cuda")
    dets_host = torch.empty((B, max_out, 8), dtype=torch.int32).pin_memory()
    cnt_host = torch.empty((B,), dtype=torch.int32).pin_memory()
    x_host2 = torch.from_numpy(synth.images(B, S, S, seed=2000 + rank)).pin_memory()
    dets_host2 = torch.empty((B, max_out, 8), dtype=torch.int32).pin_memory()
    cnt_host2 = torch.empty((B,), dtype=torch.int32).pin_memory()
    hw_host = np.array([[S, S]] * B, np.int32)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(args.warmup):
        eng.detect(x_dev, pp=pp, image_hw=hw_dev, max_out=max_out, dets=dets_dev, counts=cnt_dev, sync=False)
    eng.sync()
    barrier()
    sampler.mark()
    launches0 = eng.launch_count
    eng.timer_start()
    for _ in range(args.steps):
        eng.detect(x_dev, pp=pp, image_hw=hw_dev, max_out=max_out, dets=dets_dev, counts=cnt_dev, sync=False)
    ms_total = eng.timer_stop()
    launches = eng.launch_count - launches0
    # the timed region can be shorter than one nvidia-smi period: keep the same load running (untimed) until a few samples exist
    t_tail = time.perf_counter()
    while len(sampler.rows) - sampler.first < 5 and time.perf_counter() - t_tail < 3.0:
        for _ in range(5):
            eng.detect(x_dev, pp=pp, image_hw=hw_dev, max_out=max_out, dets=dets_dev, counts=cnt_dev, sync=False)
        eng.sync()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- kernel breakdown for the roofline (conv stack = dominant kernel), CUDA events on the handle's stream
    fwd, post = [], []
    for _ in range(min(args.steps, 10)):
        eng.detect(x_dev, pp=pp, image_hw=hw_dev, max_out=max_out, dets=dets_dev, counts=cnt_dev, sync=True)
        a, b = eng.last_timing()
        fwd.append(a); post.append(b)
    fwd_ms, post_ms = float(np.mean(fwd)), float(np.mean(post))
    n_conv = len(eng.layer_infos())
    l0 = eng.launch_count
    eng.forward(x_dev, want_outputs=False)
    n_fwd_launches = eng.launch_count - l0                           # stem + conv_igemm_kernel + conv_chain_kernel launches of one forward
    flops_step = 2.0 * eng.macs_per_image() * B                      # algorithmic, un-padded
    peaks = _peaks()
    achieved = flops_step / (fwd_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": f"conv stack of one forward: {n_conv} conv layers in {n_fwd_launches} launches (stem_strip_kernel, conv_igemm_kernel, conv_chain_kernel)", "achieved": achieved,
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                "frac_of_burst_peak": achieved / peaks["tf_burst"], "peak_source": peaks["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                "algorithmic_flops_per_step": flops_step, "avg_launch_ms": fwd_ms / max(1, n_fwd_launches), "forward_ms": fwd_ms,
                "postprocess_ms": post_ms, "share_of_step": fwd_ms / (fwd_ms + post_ms), "traffic": None}
    prof = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof))
            # dram__bytes_read.sum + dram__bytes_write.sum of the conv launches (ncu, profiles/conv_traffic.json), per launch like `achieved`
            roofline["traffic"] = tr.get("dram_bytes_per_launch")
            roofline["traffic_unit"] = "bytes per launch (DRAM read + write, ncu; average over the conv launches of one step)"
            roofline["traffic_per_step"] = tr.get("dram_bytes_per_step")
            roofline["algorithmic_bytes_per_step_unfused"] = tr.get("algorithmic_bytes_per_step_unfused")
        except Exception:
            pass

    # ---- end to end through the public API with HOST buffers (pinned): every step copies its own images H2D and its
    # detections D2H inside the timed region.  Steps are issued with the asynchronous entry point (a serving loop
    # feeding alternating input buffers), so the copy of step i+1 overlaps the compute of step i; `serial_*` is the
    # same measurement with a host synchronisation after every step (single-request latency).
    xs = (x_host, x_host2); outs_h = ((dets_host, cnt_host), (dets_host2, cnt_host2))
    for i in range(3):
        eng.detect(xs[i % 2], pp=pp, image_hw=hw_host, max_out=max_out, dets=outs_h[i % 2][0], counts=outs_h[i % 2][1], sync=True)
    barrier()
    eng.timer_start()
    for i in range(args.steps):
        eng.detect(xs[i % 2], pp=pp, image_hw=hw_host, max_out=max_out, dets=outs_h[i % 2][0], counts=outs_h[i % 2][1], sync=True)
    serial_ms = max_over_ranks(eng.timer_stop())
    barrier()
    t0 = time.perf_counter()
    eng.timer_start()
    for i in range(args.steps):
        eng.detect(xs[i % 2], pp=pp, image_hw=hw_host, max_out=max_out, dets=outs_h[i % 2][0], counts=outs_h[i % 2][1], sync=False)
    e2e_ms = max_over_ranks(eng.timer_stop())
    e2e_wall = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e = {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4 + hw_host.nbytes),
           "d2h_bytes_per_step": int(dets_host.numel() * 4 + cnt_host.numel() * 4), "ms_per_step": e2e_ms / args.steps,
           "wall_ms_per_step": e2e_wall / args.steps, "serial_ms_per_step": serial_ms / args.steps,
           "serial_value": world * B * args.steps / (serial_ms * 1e-3),
           "api": "Engine.detect(sync=False) -> fvy_detect_async (pinned host buffers, copies overlapped with the previous step's compute)"}
    kept = int(cnt_host.sum().item())
    # the same loop fed uint8 frames (FVY_U8: the device forms float32(pixel / 255) itself): a quarter of the host-to-device bytes, which is
    # what the end-to-end figure hangs on when several ranks share the host (reported next to the float32 figure, never instead of it)
    try:
        x8 = [torch.empty(tuple(x_host.shape), dtype=torch.uint8).pin_memory() for _ in range(2)]
        for t in x8:
            t.copy_((x_host * 255.0).round().clamp_(0, 255).to(torch.uint8))
        for i in range(3):
            eng.detect(x8[i % 2], pp=pp, image_hw=hw_host, max_out=max_out, dets=outs_h[i % 2][0], counts=outs_h[i % 2][1], sync=True)
        barrier()
        eng.timer_start()
        for i in range(args.steps):
            eng.detect(x8[i % 2], pp=pp, image_hw=hw_host, max_out=max_out, dets=outs_h[i % 2][0], counts=outs_h[i % 2][1], sync=False)
        u8_ms = max_over_ranks(eng.timer_stop())
        e2e["uint8_frames"] = {"value": world * B * args.steps / (u8_ms * 1e-3), "h2d_bytes_per_step": int(x_host.numel() + hw_host.nbytes),
                               "ms_per_step": u8_ms / args.steps}
    except Exception as exc:      # an additive figure: never fail the bench over it
        e2e["uint8_frames"] = {"error": str(exc)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_reference_pass(1, S, 0, threads)
        n = args.ref_images * 3
        dt, _ = cpu_reference_pass(n, S, 1, threads)
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n} images of the same workload (torch-CPU fp32 restatement of the Keras graph + C restatement of decode/NMS), {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC.replace("@416", f"@{args.size}"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"YOLOv3 face detector (nb_class=1, 18-ch heads) inference batch {B} @{S}x{S} per GPU: forward+decode_netout+correct_yolo_boxes+do_nms",
                           "net": S, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world} (batch sharded, no collective)",
                           "weights": "random-init (Keras default: glorot-uniform, identity BN), seed 0",
                           "obj_thresh": 0.5, "nms_thresh": 0.45, "anchor_mask": "reference (yolov3_detect.py:354-362)",
                           "l2": "no explicit flush: each step streams ~3.4 GB of activations (>> 126 MB L2) between reuses of the input"},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
                "gpu_launches_per_step": launches / args.steps, "clocks": clocks, "kept_boxes_last_step": kept}
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
