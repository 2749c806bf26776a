"""GPU bring-up diagnostics (run under gpurun).  Prints a layer-by-layer parity table of the CUDA
conv stack against the oracle's bf16-emulated forward, head parity against the fp32 oracle, and
decode / NMS parity against the C oracle.  Test infrastructure: may import oracle/.

usage: python tools/gpu_check.py [--size 416] [--batch 2] [--init bn_exercising] [--profile-batch 40]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from face_vijnana_yolov3_b200 import _lib as L, arch, synth   # noqa: E402
from face_vijnana_yolov3_b200.engine import Engine, post_params   # noqa: E402
from oracle import darknet_ref as D, postproc as P   # noqa: E402


def fmt_stages(v):
    a, b, res, pair, slab = v % 100, v // 100 % 100, v // 10000 % 10, v // 100000 % 10, v // 1000000
    return f"A{a}/B{b}{'R' if res else ''}{'P' if pair else ''}{'S' if slab else ''}"


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def check_forward(size, batch, init, head, tile_n_max, out):
    specs = arch.table(head, 1)
    stream = synth.darknet_stream(specs, 0, init)
    x = synth.images(batch, size, size, 0)
    eng = Engine(size, size, head=head, nb_class=1, max_batch=batch, tile_n_max=tile_n_max)
    eng.load_weights(stream)
    t0 = time.time()
    outs = eng.forward(x)
    print(f"forward ok in {time.time()-t0:.3f}s; device fwd ms={eng.last_timing()[0]:.3f}", flush=True)
    taps_emu = {}
    emu = D.forward(stream, x, 1, fd6=(head == L.HEAD_FD6), emulate_bf16=True, taps=taps_emu)
    ref = D.forward(stream, x, 1, fd6=(head == L.HEAD_FD6))
    infos = eng.layer_infos()
    rows = []
    bad = 0
    for li, info in enumerate(infos):
        got = eng.layer_output(li, batch)
        exp = taps_emu[info["idx"]].permute(0, 2, 3, 1).numpy()
        r = rel_l2(got, exp)
        flag = "" if r < 3e-2 else "  <-- BAD"
        bad += r >= 3e-2
        rows.append((info["idx"], r))
        print(f"  layer {li:2d} conv_{info['idx']:<4d} {info['cin']:4d}->{info['cout']:4d} k{info['k']} s{info['stride']} "
              f"{info['H']:3d}x{info['W']:<3d} tileN={info['tile_n']:3d} tileK={info['tile_k']} st={fmt_stages(info['stages'])} grid={info['grid']:4d} "
              f"tiles={info['tiles']:6d} relL2(vs bf16 emu)={r:.3e}{flag}", flush=True)
    if head == L.HEAD_FD6:
        outs_l, emu_l, ref_l = [outs[0]], [emu], [ref]
    else:
        outs_l, emu_l, ref_l = outs, emu, ref
    for i, (o, e, r) in enumerate(zip(outs_l, emu_l, ref_l)):
        print(f"  head {i}: relL2 vs fp32 oracle={rel_l2(o, r):.4e}  vs bf16 emu={rel_l2(o, e):.4e}  (emu vs oracle {rel_l2(e, r):.4e})", flush=True)
    out[f"forward_{head}_{size}_{init}"] = dict(bad_layers=int(bad), heads_vs_oracle=[rel_l2(o, r) for o, r in zip(outs_l, ref_l)],
                                               heads_vs_emu=[rel_l2(o, e) for o, e in zip(outs_l, emu_l)])
    eng.close()
    return bad == 0


def check_post(out):
    B = 3
    outs = synth.head_logits(B, 416, 416, 1, seed=3)
    eng = Engine(416, 416, head=L.HEAD_NONE, nb_class=1, max_batch=B)
    hw = np.array([[360, 640], [416, 416], [500, 375]], np.int32)
    for arith in (L.ARITH_F64, L.ARITH_F32):
        pp = post_params(0.5, 0.45, arith=arith)
        d = eng.decode(outs, pp=pp, image_hw=hw)
        ok_all = True
        for b in range(B):
            o = P.decode_image([t[b] for t in outs], obj_thresh=0.5, arith=arith)
            ib = P.correct_yolo_boxes(o["box"], hw[b, 0], hw[b, 1], 416, 416, arith)
            n = int(d["counts"][b])
            ok = n == len(o["cell"]) and np.array_equal(d["nbox"][b, :n], o["box"]) and np.array_equal(d["ibox"][b, :n], ib) \
                and np.array_equal(d["objness"][b, :n], o["objness"]) and np.array_equal(d["classes"][b, :n], o["classes"])
            ok_all &= bool(ok)
            print(f"  decode arith={arith} img{b}: n={n} oracle_n={len(o['cell'])} exact={ok}", flush=True)
            # nms on the oracle's candidates
            cls_ref = P.do_nms(ib, o["classes"], 0.45)
            S = eng.cap
            ibp = np.zeros((1, S, 4), np.int32); ibp[0, :n] = ib
            clp = np.zeros((1, S, 1), np.float32); clp[0, :n] = o["classes"]
            cls_gpu, kept, kc = eng.nms(ibp, clp, np.array([n], np.int32), 0.45)
            okn = np.array_equal(cls_gpu[0, :n], cls_ref) and np.array_equal(kept[0, :kc[0]], np.nonzero(cls_ref[:, 0] > 0)[0])
            ok_all &= bool(okn)
            print(f"  nms    arith={arith} img{b}: kept={int(kc[0])} oracle_kept={int((cls_ref[:,0]>0).sum())} exact={okn}", flush=True)
        out[f"post_arith{arith}"] = bool(ok_all)
    eng.close()


def profile(size, batch, tile_n_max, out):
    specs = arch.table(L.HEAD_YOLO3, 1)
    stream = synth.darknet_stream(specs, 0, synth.INIT_KERAS_DEFAULT)
    eng = Engine(size, size, nb_class=1, max_batch=batch, tile_n_max=tile_n_max)
    eng.load_weights(stream)
    x = synth.images(batch, size, size, 1)
    import torch
    xd = torch.from_numpy(x).cuda()
    for _ in range(3):
        eng.forward(xd, want_outputs=False)
    ts = []
    for _ in range(10):
        eng.forward(xd, want_outputs=False)
        ts.append(eng.last_timing()[0])
    ms = float(np.median(ts))
    flops = 2.0 * eng.macs_per_image() * batch
    print(f"forward batch {batch} @{size}: median {ms:.3f} ms  => {batch/ms*1e3:.0f} img/s, {flops/ms/1e9:.1f} TFLOP/s "
          f"({flops/ms/1e9/1685.7*100:.1f}% of measured burst bf16 peak) tile_n_max={tile_n_max}", flush=True)
    lay = eng.profile_layers(batch, 5)
    infos = eng.layer_infos()
    tot = 0.0
    for info, t in zip(infos, lay):
        macs = info["H"] * info["W"] * info["cout"] * info["k"] ** 2 * info["cin"] * batch
        tf = 2 * macs / (t * 1e-3) / 1e12
        tot += t
        print(f"    conv_{info['idx']:<4d} {info['cin']:4d}->{info['cout']:4d} k{info['k']} s{info['stride']} {info['H']:3d}^2 tileN={info['tile_n']:3d} "
              f"{t*1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
    print(f"    sum of isolated layers: {tot:.3f} ms", flush=True)
    out[f"profile_{size}_b{batch}_tn{tile_n_max}"] = dict(ms=ms, img_s=batch / ms * 1e3, tflops=flops / ms / 1e9,
                                                        layers_us=[float(t * 1e3) for t in lay])
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--profile-batch", type=int, default=40)
    ap.add_argument("--skip-forward", action="store_true")
    ap.add_argument("--skip-post", action="store_true")
    ap.add_argument("--tile-n", type=int, nargs="*", default=[128, 256])
    args = ap.parse_args()
    out = {}
    os.makedirs("gpurun_out", exist_ok=True)
    print(L.load().fvy_version().decode(), flush=True)
    steps = []
    if not args.skip_post:
        steps.append(("post", lambda: check_post(out)))
    if not args.skip_forward:
        steps.append(("fwd yolo3 bn", lambda: check_forward(args.size, args.batch, synth.INIT_BN_EXERCISING, L.HEAD_YOLO3, 0, out)))
        steps.append(("fwd yolo3 keras tn256", lambda: check_forward(args.size, args.batch, synth.INIT_KERAS_DEFAULT, L.HEAD_YOLO3, 256, out)))
        steps.append(("fwd fd6", lambda: check_forward(416, args.batch, synth.INIT_BN_EXERCISING, L.HEAD_FD6, 0, out)))
    if args.profile_batch > 0:
        for tn in args.tile_n:
            steps.append((f"profile tn{tn}", lambda tn=tn: profile(args.size, args.profile_batch, tn, out)))
    for name, fn in steps:
        print(f"== {name}", flush=True)
        try:
            fn()
        except Exception as e:   # keep going: every step is independent evidence
            traceback.print_exc()
            out[f"error_{name}"] = repr(e)
            if "CUDA" in repr(e) or "cuda" in repr(e):
                print("CUDA error: stopping (context is likely dead)", flush=True)
                break
    with open("gpurun_out/gpu_check.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if not k.startswith("profile")}, indent=1))


if __name__ == "__main__":
    main()
