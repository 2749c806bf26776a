#!/usr/bin/env python
"""Headline benchmark: images/sec @416 for forward + decode + NMS (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...   (the reference's CPU path, oracle port)
  python bench.py --config {headline,608x320,stress,train}  (BASELINE.json configs[1] (default) / [2] / [4] / [3])

One step = one pass of the hot path (Darknet-53/YOLOv3 forward -> decode_netout/correct_yolo_boxes
-> do_nms) over one batch of 40 synthetic 416x416 images per GPU (BASELINE.json configs[1]).
`value` is timed with inputs resident in HBM; `e2e` goes through the public API with host buffers.
Every roofline figure is computed from CUDA events recorded INSIDE the timed region (fvy_timer_breakdown).
"""
from __future__ import annotations

import argparse
import hashlib
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


def _lib_sha16():
    try:
        return hashlib.sha256(open(os.path.join(ROOT, "face_vijnana_yolov3_b200", "libfvy.so"), "rb").read()).hexdigest()[:16]
    except OSError:
        return None


def _src_sha16():
    """Identity of the library's SOURCES (csrc/ + include/fvy.h, sorted by name).  nvcc's output is not byte-reproducible (two builds of
    the same tree differ), so evidence captured on one build of these sources is matched to another build of them by this hash."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "face_vijnana_yolov3_b200", "csrc")
    try:
        for name in sorted(os.listdir(d)):
            if name.endswith((".cu", ".cuh", ".h", ".inl")):
                h.update(name.encode()); h.update(open(os.path.join(d, name), "rb").read())
        h.update(open(os.path.join(ROOT, "include", "fvy.h"), "rb").read())
    except OSError:
        return None
    return h.hexdigest()[:16]


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

    def summary(self, first=None, last=None):
        rows = self.rows[self.first if first is None else first:last]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self.summary()


def pin_to_gpu_numa(index):
    """Best effort: run this rank's host threads (and first-touch its pinned buffers) on the NUMA node the GPU hangs off, so
    eight ranks feeding eight GPUs from one host do not all pull their images across the socket interconnect."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port timed on the host cores (the reference is pure Python over Keras/TF, which
# cannot be installed here: DESIGN.md section 4)
# ----------------------------------------------------------------------------------------------------------------------
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


def cpu_stress_pass(size):
    """One image of BASELINE configs[4] through the C oracle (decode + correct + do_nms).  Returns (seconds, candidates)."""
    from face_vijnana_yolov3_b200 import synth
    from oracle import postproc as P
    outs = synth.head_logits(1, size, size, 1, seed=4, crowd=True, obj_bias=6.0)
    t0 = time.perf_counter()
    dd = P.decode_image([o[0] for o in outs], anchor_masks=P.ALL_ANCHOR_MASK, obj_thresh=0.5, net_h=size, net_w=size)
    ib = P.correct_yolo_boxes(dd["box"], size, size, size, size)
    P.do_nms(ib, dd["classes"], 0.5)
    return time.perf_counter() - t0, len(ib)


def run_reference(args, rank):
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = args.ref_images
    if args.config == "stress":
        t, n = 0.0, 0
        for _ in range(max(1, args.steps)):
            dt, n = cpu_stress_pass(args.size)
            t += dt
        v = max(1, args.steps) / t
        what = f"1 image per step x {max(1, args.steps)} steps through the C oracle (decode + correct_yolo_boxes + do_nms, {n} candidates), 1 core"
        threads = 1
        metric, workload = _stress_names(args)
    elif args.config == "train":
        print(json.dumps({"impl": "reference", "unavailable": "the reference's training step is Keras/TF (not installable here); "
                          "bench.py --config train reports the cuDNN/autograd step as its stated baseline"}), file=_OUT, flush=True)
        return 0
    else:
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_reference_pass(1, args.size, 0, threads)
        t = 0.0
        for s in range(args.steps):
            dt, _ = cpu_reference_pass(sample, args.size, s, threads)
            t += dt
        v = sample * args.steps / t
        what = (f"{sample} of the {args.batch} images per step x {args.steps} steps; torch-CPU fp32 restatement of the Keras graph "
                "+ C restatement of decode/NMS (Keras/TF not installable; reference is pure Python)")
        metric = METRIC.replace("@416", f"@{args.size}")
        workload = f"YOLOv3 face detector (nb_class=1) inference batch {args.batch} @{args.size}x{args.size}: forward+decode_netout+do_nms"
    line = {"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "net": args.size, "batch_per_gpu": args.batch, "sample_images_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": what},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL_DEBUG=INFO lines, version banners), so file
    descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the real stdout."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_OUT = sys.stdout


class Dist:
    """Rank plumbing shared by every config: barrier + synchronize on both sides of a timed region, max over ranks."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback for the CUDA hot path")
        torch.cuda.set_device(self.local)
        self.numa = pin_to_gpu_numa(self.local)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# configs[1] (headline) and configs[2] (608x320, the global batch sharded over the ranks)
# ----------------------------------------------------------------------------------------------------------------------
def post_algorithmic_bytes(size, n_cands, n_kept_boxes):
    """decode reads every head logit once and writes 28 B per candidate; NMS reads n sorted records (24 B) and writes +
    reads the n^2/8-byte suppression bitmask (upper triangle of 64-bit words), + 32 B per detection record written."""
    n_all = sum(3 * (size // s) * (size // s) for s in (32, 16, 8))
    words = (n_cands + 63) // 64
    return n_all * 6 * 4 + 28 * n_cands + 24 * n_cands + 2 * n_cands * words * 8 + 32 * n_kept_boxes


def run_detect(args, D):
    import torch
    from face_vijnana_yolov3_b200 import _lib as L, arch, synth
    from face_vijnana_yolov3_b200.engine import Engine, post_params
    rank, world, local_rank = D.rank, D.world, D.local
    strong = args.config == "608x320"
    if strong:
        if 320 % world:
            raise SystemExit("--config 608x320 needs a world size that divides 320")
        args.size, args.batch = 608, 320 // world
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

    def dev_step():
        eng.detect(x_dev, pp=pp, image_hw=hw_dev, max_out=max_out, dets=dets_dev, counts=cnt_dev, sync=False)

    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(args.warmup):
        dev_step()
    eng.sync()
    D.barrier()
    sampler.mark()
    launches0 = eng.launch_count
    eng.timer_start()
    for _ in range(args.steps):
        dev_step()
    ms_total = eng.timer_stop()
    fwd_ms, post_ms, n_timed = eng.timer_breakdown()         # CUDA events of these very steps
    launches = eng.launch_count - launches0
    eng.sync()                                                # deferred errors of the asynchronous calls surface here
    # the timed region can be shorter than one nvidia-smi period: keep the same load running (untimed) until a few samples exist
    t_tail = time.perf_counter()
    while len(sampler.rows) - sampler.first < 5 and time.perf_counter() - t_tail < 3.0:
        for _ in range(5):
            dev_step()
        eng.sync()
    D.barrier()
    clocks = sampler.summary()
    ms_total = D.max(ms_total)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- the same loop for >= 1 s: what the GPU sustains under its power cap (the 20-step region above is a burst)
    sustained = None
    if args.sustained_s > 0:
        n_sus = max(args.steps, int(np.ceil(args.sustained_s * 1e3 / ms_step)))
        D.barrier()
        s_first = len(sampler.rows)
        eng.timer_start()
        for _ in range(n_sus):
            dev_step()
        sus_ms = D.max(eng.timer_stop())
        sus_fwd, sus_post, _ = eng.timer_breakdown()
        eng.sync()
        sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "value": world * B * n_sus / (sus_ms * 1e-3), "unit": UNIT,
                     "ms_per_step": sus_ms / n_sus, "forward_ms": sus_fwd, "postprocess_ms": sus_post,
                     "clocks": sampler.summary(first=s_first)}
    sampler.stop()

    # ---- the two halves timed ALONE (synchronous calls: no overlap between the post-processing of one step and the forward of
    # the next), for comparison with the in-region figures above
    fa, pa = [], []
    for _ in range(min(args.steps, 10)):
        eng.detect(x_dev, pp=pp, image_hw=hw_dev, max_out=max_out, dets=dets_dev, counts=cnt_dev, sync=True)
        a_, b_ = eng.last_timing()
        fa.append(a_); pa.append(b_)
    fwd_alone, post_alone = float(np.median(fa)), float(np.median(pa))

    # ---- rooflines.  Conv stack (tensor bound): algorithmic FLOPs / the forward's CUDA-event time inside the timed steps.
    n_conv = len(eng.layer_infos())
    l0 = eng.launch_count
    eng.forward(x_dev, want_outputs=False)
    n_fwd_launches = eng.launch_count - l0                           # stem + conv_igemm_kernel + conv_chain_kernel launches of one forward
    flops_step = 2.0 * eng.macs_per_image() * B                      # algorithmic, un-padded
    peaks = _peaks()
    achieved = flops_step / (fwd_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": f"conv stack of one forward: {n_conv} conv layers in {n_fwd_launches} launches (stem_strip_kernel, conv_igemm_kernel, conv_chain_kernel)",
                "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
                "peak_source": peaks["source"] + f", burst bf16 figure (timed region {ms_total:.0f} ms)",
                "frac_of_sustained_peak": achieved / peaks["tf_sustained"],
                "algorithmic_flops_per_step": flops_step, "avg_launch_ms": fwd_ms / max(1, n_fwd_launches), "forward_ms": fwd_ms,
                "forward_ms_source": f"CUDA events around the forward of each of the {n_timed} timed steps (fvy_timer_breakdown)",
                "postprocess_ms": post_ms, "share_of_step": fwd_ms / max(ms_step, 1e-9),
                "forward_ms_alone": fwd_alone, "postprocess_ms_alone": post_alone,
                "note": "in the timed (asynchronous) steps the post-processing of step i runs on a low-priority stream beside the forward of step i+1: "
                        "forward_ms includes that interference and postprocess_ms is stretched by it; *_alone are the same kernels in synchronous calls",
                "traffic": None}
    if sustained:
        a_s = flops_step / (sustained["forward_ms"] * 1e-3) / 1e12
        sustained["roofline"] = {"achieved": a_s, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": a_s / peaks["tf_sustained"],
                                 "peak_source": peaks["source"] + ", sustained bf16 figure"}
    prof = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof))
            # dram__bytes_read.sum + dram__bytes_write.sum of the conv launches (ncu, tools/evidence_r02.sh), per launch like `achieved`
            roofline["traffic"] = tr.get("dram_bytes_per_launch")
            roofline["traffic_unit"] = "bytes per launch (DRAM read + write, ncu; average over the conv launches of one step)"
            roofline["traffic_per_step"] = tr.get("dram_bytes_per_step")
            roofline["algorithmic_bytes_per_step_unfused"] = tr.get("algorithmic_bytes_per_step_unfused")
            roofline["traffic_build"] = tr.get("libfvy_src_sha16") or tr.get("libfvy_sha16")
            same = (tr.get("libfvy_src_sha16") == _src_sha16()) if tr.get("libfvy_src_sha16") else (tr.get("libfvy_sha16") == _lib_sha16())
            roofline["traffic_is_this_build"] = bool(same) and (tr.get("batch"), tr.get("net")) == (B, S)
        except Exception:
            pass
    # Decode + NMS (HBM bound): algorithmic bytes per SURVEY 8(d) / the post-processing's CUDA-event time inside the timed steps
    d = eng.decode(batch=B, pp=pp, image_hw=hw_host, want_nbox=False)
    n_c = d["counts"].astype(np.int64)
    cnt_now = cnt_dev.cpu().numpy().astype(np.int64)
    post_bytes = float(sum(post_algorithmic_bytes(S, int(n_c[b]), int(cnt_now[b])) for b in range(B)))
    post_ach = post_bytes / (post_ms * 1e-3) / 1e9
    roofline_post = {"bound": "hbm", "kernel": "decode_yolo_kernel + sort_scores_kernel + nms_mask_kernel + nms_sweep_kernel + assemble_yolo_kernel",
                     "achieved": post_ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": post_ach / peaks["hbm"], "peak_source": peaks["source"],
                     "postprocess_ms": post_ms, "postprocess_ms_alone": post_alone, "achieved_alone": post_bytes / (post_alone * 1e-3) / 1e9,
                     "algorithmic_bytes_per_step": post_bytes, "candidates_per_image_mean": float(n_c.mean()),
                     "kept_per_image_mean": float(cnt_now.mean()),
                     "note": "launch/latency-bound at ~2k candidates per image, not bandwidth-bound (DESIGN.md 3.3); BASELINE configs[4] "
                             "(--config stress) is the size the HBM roofline is meaningful at"}

    # ---- end to end through the public API with HOST buffers (pinned): every step copies its own images H2D and its
    # detections D2H inside the timed region.  Steps are issued with the asynchronous entry point (a serving loop
    # feeding alternating input buffers), so the copy of step i+1 overlaps the compute of step i; `serial_*` is the
    # same measurement with a host synchronisation after every step (single-request latency).
    # The frames a pipeline holds are uint8 (imread): the headline `e2e.value` ships them as they are (FVY_U8: the device forms
    # float32(pixel / 255) itself, bit-identical logits, tests/test_gpu_parity.py::test_uint8_images_equal_image_over_255); the
    # float32 form the reference's `detect(image / 255)` call site passes is measured beside it (`float32_frames`).
    outs_h = ((dets_host, cnt_host), (dets_host2, cnt_host2))

    def host_loop(xs, sync):
        for i in range(3):
            eng.detect(xs[i % 2], pp=pp, image_hw=hw_host, max_out=max_out, dets=outs_h[i % 2][0], counts=outs_h[i % 2][1], sync=True)
        D.barrier()
        t0 = time.perf_counter()
        eng.timer_start()
        for i in range(args.steps):
            eng.detect(xs[i % 2], pp=pp, image_hw=hw_host, max_out=max_out, dets=outs_h[i % 2][0], counts=outs_h[i % 2][1], sync=sync)
        ms = D.max(eng.timer_stop())
        wall = D.max((time.perf_counter() - t0) * 1e3)
        eng.sync()
        return ms, wall

    x8 = [torch.empty(tuple(x_host.shape), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for t, src in zip(x8, (x_host, x_host2)):
        t.copy_((src * 255.0).round().clamp_(0, 255).to(torch.uint8))
    d2h = int(dets_host.numel() * 4 + cnt_host.numel() * 4)
    u8_serial, _ = host_loop(x8, True)
    u8_ms, u8_wall = host_loop(x8, False)
    f32_serial, _ = host_loop((x_host, x_host2), True)
    f32_ms, f32_wall = host_loop((x_host, x_host2), False)
    e2e = {"value": world * B * args.steps / (u8_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x8[0].numel() + hw_host.nbytes),
           "d2h_bytes_per_step": d2h, "ms_per_step": u8_ms / args.steps, "wall_ms_per_step": u8_wall / args.steps,
           "serial_ms_per_step": u8_serial / args.steps, "serial_value": world * B * args.steps / (u8_serial * 1e-3),
           "input": "uint8 frames (B, H, W, 3) as imread returns them; the device forms float32(pixel / 255.0)",
           "api": "Engine.detect(sync=False) -> fvy_detect_async (pinned host buffers, copies overlapped with the previous step's compute)",
           "float32_frames": {"value": world * B * args.steps / (f32_ms * 1e-3), "h2d_bytes_per_step": int(x_host.numel() * 4 + hw_host.nbytes),
                              "ms_per_step": f32_ms / args.steps, "wall_ms_per_step": f32_wall / args.steps,
                              "serial_value": world * B * args.steps / (f32_serial * 1e-3),
                              "input": "float32 image / 255 (what the reference's detect() call site passes)"},
           "numa": D.numa}
    kept = int(cnt_host.sum().item())

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
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": (f"BASELINE configs[2]: YOLOv3 face detector inference batch 320 @608x608 sharded over {world} GPU(s), {B} images per GPU per step"
                                        if strong else
                                        f"YOLOv3 face detector (nb_class=1, 18-ch heads) inference batch {B} @{S}x{S} per GPU: forward+decode_netout+correct_yolo_boxes+do_nms"),
                           "net": S, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world} (batch sharded, no collective)",
                           "weights": "random-init (Keras default: glorot-uniform, identity BN), seed 0",
                           "obj_thresh": 0.5, "nms_thresh": 0.45, "anchor_mask": "reference (yolov3_detect.py:354-362)",
                           "l2": "no explicit flush: each step streams ~3.4 GB of activations (>> 126 MB L2) between reuses of the input"},
                "roofline": roofline, "roofline_post": roofline_post, "sustained": sustained, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps, "clocks": clocks, "kept_boxes_last_step": kept,
                "libfvy_sha16": _lib_sha16(), "libfvy_src_sha16": _src_sha16()}
        print(json.dumps(line), file=_OUT, flush=True)
    eng.close()
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# configs[4]: decode / NMS stress
# ----------------------------------------------------------------------------------------------------------------------
def _stress_names(args):
    n = sum(3 * (args.size // s) * (args.size // s) for s in (32, 16, 8))
    return ("images/sec (decode+NMS stress, ~10k candidates/image)",
            f"BASELINE configs[4]: crowd logits, all 9 anchors, {n} candidates/image, batch {args.batch} @{args.size}, nms_iou_th 0.5, num_cands 60")


def run_stress(args, D):
    """Head logits resident in HBM; one step = decode_netout + correct_yolo_boxes + do_nms + selection for the batch
    (fvy_postprocess); e2e = the same with the logits in pinned host memory.  Roofline: SURVEY 8(d) bytes / the step's CUDA-event time."""
    import torch
    from face_vijnana_yolov3_b200 import _lib as L, synth
    from face_vijnana_yolov3_b200.engine import DET_DTYPE, Engine, post_params
    B, S = args.batch, args.size
    outs = synth.head_logits(B, S, S, 1, seed=4 + D.rank, crowd=True, obj_bias=6.0)
    host = [torch.from_numpy(np.ascontiguousarray(o)).pin_memory() for o in outs]
    dev = [t.cuda() for t in host]
    eng = Engine(S, S, head=L.HEAD_NONE, nb_class=1, max_batch=B, device=D.local)
    pp = post_params(0.5, 0.5, anchor_mask=L.ANCHOR_MASK_ALL, num_cands=60)
    hw = np.array([[S, S]] * B, np.int32)
    sampler = ClockSampler(D.local)
    sampler.start()
    for _ in range(args.warmup):
        dets, counts = eng.postprocess(dev, pp=pp, image_hw=hw, max_out=60)
    D.barrier()
    sampler.mark()
    l0 = eng.launch_count
    eng.timer_start()
    for _ in range(args.steps):
        dets, counts = eng.postprocess(dev, pp=pp, image_hw=hw, max_out=60)
    ms = D.max(eng.timer_stop()) / args.steps
    launches = eng.launch_count - l0
    D.barrier()
    eng.timer_start()
    for _ in range(args.steps):
        dets, counts = eng.postprocess(host, pp=pp, image_hw=hw, max_out=60)
    e2e_ms = D.max(eng.timer_stop()) / args.steps
    t_tail = time.perf_counter()
    while len(sampler.rows) - sampler.first < 5 and time.perf_counter() - t_tail < 3.0:
        eng.postprocess(dev, pp=pp, image_hw=hw, max_out=60)
    clocks = sampler.stop()
    n = sum(3 * (S // s) * (S // s) for s in (32, 16, 8))      # every cell x anchor passes (tests/test_gpu_parity.py::test_nms_stress_crowd_10k)
    alg = float(B * post_algorithmic_bytes(S, n, 60))
    peaks = _peaks()
    achieved = alg / (ms * 1e-3) / 1e9
    metric, workload = _stress_names(args)
    line = {"metric": metric, "value": D.world * B / (ms * 1e-3), "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 / f64 IoU, f32 scores", "data": "synthetic",
            "config": {"workload": workload, "net": S, "batch_per_gpu": B, "pairs_per_image": n * (n - 1) // 2,
                       "l2": "the per-image suppression bitmask (14.3 MB x batch = 570 MB) is larger than L2"},
            "roofline": {"bound": "hbm", "kernel": "decode_yolo_kernel + sort_scores_kernel + nms_mask_kernel + nms_sweep_kernel + assemble_yolo_kernel",
                         "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"], "peak_source": peaks["source"],
                         "algorithmic_bytes_per_step": alg, "traffic": None,
                         "note": "the bitmask kernel (n^2/2 IoU tests per image) is instruction-bound at this size, the sort and the sweep are "
                                 "dependent chains per image - see profiles/ for the per-kernel ncu list"},
            "e2e": {"value": D.world * B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host) + hw.nbytes),
                    "d2h_bytes_per_step": int(B * 60 * 32 + B * 4), "ms_per_step": e2e_ms,
                    "api": "Engine.postprocess -> fvy_postprocess with the head logits in pinned host memory"},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps, "clocks": clocks,
            "kept_per_image_mean": float(np.mean(counts)), "libfvy_sha16": _lib_sha16()}
    if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
        dt, nc = cpu_stress_pass(S)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port", "candidates": int(nc),
                                "sample": f"1 image of the same workload through the C oracle (decode + correct + do_nms), {dt:.2f} s"}
    if D.rank == 0:
        print(json.dumps(line), file=_OUT, flush=True)
    eng.close()
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# configs[3]: FaceDetector training step, data-parallel with the NCCL gradient all-reduce
# ----------------------------------------------------------------------------------------------------------------------
def run_train(args, D):
    """Global batch 40 @416 split over the ranks (strong scaling: multi_gpu_model's axis-0 split), fp32 like the reference's
    Keras training (bf16 autocast reported beside it, never as the parity number).  Exposed communication = step time minus the
    time of the same step with the all-reduce switched off."""
    import torch
    from face_vijnana_yolov3_b200 import arch, synth, train as T
    world, rank, local = D.world, D.rank, D.local
    GB, S = 40, 416
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False          # the reference trains in fp32 (TF 1.13): no TF32 in the parity number
    torch.backends.cuda.matmul.allow_tf32 = False
    hps = dict(lr=1e-4, beta_1=0.99, beta_2=0.99, decay=0.0)
    stream = synth.darknet_stream(arch.fd6_table(6), 0, synth.INIT_KERAS_DEFAULT)
    images = synth.images(GB, S, S, 0)
    targets = T.synthetic_targets(GB, 1, cell_size=S // 32)
    xs, ts = T.slice_for_rank(images, targets, rank, world)
    xs, ts = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ts).pin_memory()

    def timed(tr, steps, exchange=True):
        tr.exchange = exchange
        for _ in range(args.warmup):
            loss = tr.step(xs, ts, global_batch=GB)
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = tr.step(xs, ts, global_batch=GB)
        e1.record()
        torch.cuda.synchronize()
        tr.exchange = True
        return D.max(e0.elapsed_time(e1) / steps), loss

    sampler = ClockSampler(local)
    sampler.start()
    tr_base = T.DataParallelTrainer(hps, device=f"cuda:{local}", stream=stream, bucket_mb=args.bucket_mb, autocast_bf16=False, fvy_bn=False)
    ms_cudnn, _ = timed(tr_base, args.steps)       # the stated baseline: every layer on torch autograd / cuDNN
    del tr_base
    torch.cuda.empty_cache()
    tr = T.DataParallelTrainer(hps, device=f"cuda:{local}", stream=stream, bucket_mb=args.bucket_mb, autocast_bf16=False)
    sampler.mark()
    ms, loss = timed(tr, args.steps)
    ms_noex = timed(tr, args.steps, exchange=False)[0] if world > 1 else ms
    ar_ms = None
    if world > 1:
        for _ in range(2):
            tr.exchange_all()
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            tr.exchange_all()
        e1.record()
        torch.cuda.synchronize()
        ar_ms = D.max(e0.elapsed_time(e1) / 5)
    nbytes = tr.n_params * 4
    n_buckets = len(tr.buckets)
    del tr
    torch.cuda.empty_cache()
    tr16 = T.DataParallelTrainer(hps, device=f"cuda:{local}", stream=stream, bucket_mb=args.bucket_mb, autocast_bf16=True)
    ms16, _ = timed(tr16, args.steps)
    del tr16
    torch.cuda.empty_cache()
    # the same fp32 step with the input gradients of the eligible stride-1 convolutions on this repo's tcgen05 kernel (bf16 operands)
    tr_tc = T.DataParallelTrainer(hps, device=f"cuda:{local}", stream=stream, bucket_mb=args.bucket_mb, autocast_bf16=False, fvy_dgrad=True)
    ms_tc, loss_tc = timed(tr_tc, args.steps)
    n_tc = sum(1 for c in tr_tc.model.convs.values() if T._tc_eligible(c) and c.bias is None)
    del tr_tc
    torch.cuda.empty_cache()
    # ... plus the weight gradients on conv_wgrad_kernel (mode 3), plus the forward on the tcgen05 kernel (mode 7: every stride-1 conv on this repo's kernels)
    fvy_conv = {}
    for mode, key in ((3, "fvy_conv_backward"), (7, "fvy_conv_all")):
        from face_vijnana_yolov3_b200 import conv_tc
        conv_tc.clear_cache()
        trm = T.DataParallelTrainer(hps, device=f"cuda:{local}", stream=stream, bucket_mb=args.bucket_mb, autocast_bf16=False, fvy_conv_mode=mode)
        ms_m, loss_m = timed(trm, args.steps)
        fvy_conv[key] = {"ms_per_step": ms_m, "value": GB / (ms_m * 1e-3), "loss": loss_m, "mode": mode}
        del trm
        torch.cuda.empty_cache()
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": "FaceDetector training images/sec (fwd+bwd+allreduce+Adam)", "value": GB / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "fp32 (TF32 off; the reference trains in fp32)", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[3]: FaceDetector (Darknet-53 base + 3x3x6 head) training step, global batch {GB} @{S} over {world} GPU(s), "
                                       "MSE vs synthetic 13x13x6 targets, Keras Adam(1e-4, 0.99, 0.99)", "net": S, "global_batch": GB,
                           "batch_per_gpu": GB // world if GB % world == 0 else f"{GB}/{world}", "parallelism": f"dp{world}, BatchNorm statistics per GPU, one NCCL all-reduce of the fp32 gradients"},
                "loss": loss, "params": nbytes // 4,
                "exchange": {"allreduce_bytes_per_step": nbytes if world > 1 else 0, "buckets": n_buckets, "bucket_mb": args.bucket_mb,
                             "allreduce_alone_ms": ar_ms, "bus_GBps": (2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9) if ar_ms else None,
                             "step_ms_without_exchange": ms_noex, "exposed_communication_ms": max(0.0, ms - ms_noex)},
                "baseline_all_cudnn": {"ms_per_step": ms_cudnn, "value": GB / (ms_cudnn * 1e-3),
                                       "note": "the same fp32 step with BatchNorm / LeakyReLU on torch's modules as well (everything library code)"},
                "bf16_autocast": {"ms_per_step": ms16, "value": GB / (ms16 * 1e-3), "note": "narrower arithmetic than the reference's training; not the parity number"},
                "tc_dgrad": {"ms_per_step": ms_tc, "value": GB / (ms_tc * 1e-3), "loss": loss_tc, "layers": n_tc,
                             "note": "fp32 step with dX of the stride-1 convolutions on conv_igemm_kernel (fvy_conv_run: bf16 operands, fp32 accumulation; "
                                     "forward and dW stay fp32 on cuDNN); gradient buckets within 1e-2 relative L2 of the all-fp32 step (GPU test) - opt-in, not the parity number"},
                "fvy_conv_backward": dict(fvy_conv["fvy_conv_backward"], note="tc_dgrad plus dW of the stride-1 convolutions (Cin, Cout multiples of 64) on conv_wgrad_kernel "
                                          "(fvy_conv_wgrad: pixel-dimension GEMM, warp-level bf16 MMAs, fp32 accumulation); forward fp32 on cuDNN"),
                "fvy_conv_all": dict(fvy_conv["fvy_conv_all"], note="forward, dX and dW of every stride-1 convolution on this repo's kernels (bf16 operands): compare with bf16_autocast, "
                                     "the library's step at the same operand precision"),
                "compute": "hand-written: BatchNorm (batch statistics) + LeakyReLU forward / backward for all 52 pairs (fvy_bn_leaky_train_*), Keras Adam "
                           "(fvy_adam_step), bucketed exchange + overlap, weight-stream interop; conv forward / dgrad / wgrad of the headline value: torch autograd over cuDNN "
                           "(library code); `tc_dgrad` is the step with dgrad on this repo's tcgen05 kernel - wgrad is the part of row f-1 still open",
                "roofline": None, "cpu_baseline": None, "e2e": {"value": GB / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(xs.numel() * 4 + ts.numel() * 4),
                                                               "d2h_bytes_per_step": 4, "note": "every step copies its images / targets from pinned host memory and reads the loss back"},
                "gpu_launches": None, "clocks": clocks}
        print(json.dumps(line), file=_OUT, flush=True)
    return 0


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="fvy", choices=["fvy", "reference"])
    ap.add_argument("--config", default="headline", choices=["headline", "608x320", "stress", "train"],
                    help="BASELINE.json configs[1] (default), [2] (batch 320 @608 sharded over the ranks), [4] (decode/NMS stress), [3] (training step)")
    ap.add_argument("--batch", type=int, default=40, help="images per GPU per step (BASELINE configs[1])")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--ref-images", type=int, default=2, help="images per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustained-s", type=float, default=1.0, help="length of the additional sustained timed region (0 = skip)")
    ap.add_argument("--bucket-mb", type=float, default=256.0, help="gradient bucket size: one 162.6 MB bucket reaches 608 GB/s bus bandwidth at N = 8 (three 64 MB buckets: 394)")
    ap.add_argument("--tile-n", type=int, default=0)
    args = ap.parse_args()
    dflt_steps = {"headline": 20, "608x320": 5, "stress": 20, "train": 10}[args.config]
    args.steps = dflt_steps if args.steps is None else args.steps
    args.warmup = 5 if args.warmup is None else args.warmup
    if args.warmup < 3 and args.impl == "fvy":
        args.warmup = 3
    if args.impl == "reference":
        if args.config == "608x320":
            args.size = 608
        return run_reference(args, int(os.environ.get("RANK", "0")))
    D = Dist()
    rc = {"headline": run_detect, "608x320": run_detect, "stress": run_stress, "train": run_train}[args.config](args, D)
    D.close()
    return rc


if __name__ == "__main__":
    sys.exit(main())
