"""ORACLE — test infrastructure only.  torch-CPU fp32 restatement of the reference's Keras graph.

PARITY UNPINNED at the Keras/TensorFlow boundary: Keras 2.2.4 / TF 1.13.1 (setup.py:23,
README.md:15-16) are third-party, un-vendored and not installable here, and the reference has no
tests or golden tensors for the forward pass.  What IS pinned: the topology / per-layer shapes /
parameter counts against the Keras ``summary()`` dumps the reference ships
(``analysis/face_recog_analysis.ipynb:1431-1868``: 52 convs, 23 adds, 13x13x1024, 40,620,640
params) — see ``tests/test_oracle_forward.py``.

Restates, line-aligned:
  * ``_conv_block``          src/space/yolov3_detect.py:196-215
  * ``make_yolov3_model``    src/space/yolov3_detect.py:217-311   (head width is a parameter)
  * FaceDetector base+head   src/space/face_detection.py:341-352, 405-595
  * weight stream order      src/space/yolov3_detect.py:91-119
Keras semantics restated: ZeroPadding2D(1) = symmetric 1-px zero pad; Conv2D(padding='valid') =
cross-correlation; BatchNormalization(epsilon=0.001) inference = gamma*(x-mean)/sqrt(var+eps)+beta;
LeakyReLU(alpha=0.1); UpSampling2D(2) = nearest; concatenate = channel axis, upsampled first.

``emulate_bf16=True`` is NOT the oracle: it predicts the CUDA path's numerics (BN folded in fp32,
weights rounded to bf16 once, every stored activation rounded to bf16, fp32 accumulate, fp32 heads)
and is used by tests for layer-wise debugging only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 0.001  # yolov3_detect.py:212


def _c(f, k, s, bn, leaky, idx):
    return {"filter": f, "kernel": k, "stride": s, "bnorm": bn, "leaky": leaky, "layer_idx": idx}


def graph_blocks(head_c: int):
    """The literal block list of make_yolov3_model (:221-308) as (name, input, convs, skip)."""
    T, Fa = True, False
    B = []
    B.append(("b0", "in", [_c(32, 3, 1, T, T, 0), _c(64, 3, 2, T, T, 1), _c(32, 1, 1, T, T, 2), _c(64, 3, 1, T, T, 3)], True))
    B.append(("b5", "b0", [_c(128, 3, 2, T, T, 5), _c(64, 1, 1, T, T, 6), _c(128, 3, 1, T, T, 7)], True))
    B.append(("b9", "b5", [_c(64, 1, 1, T, T, 9), _c(128, 3, 1, T, T, 10)], True))
    B.append(("b12", "b9", [_c(256, 3, 2, T, T, 12), _c(128, 1, 1, T, T, 13), _c(256, 3, 1, T, T, 14)], True))
    prev = "b12"
    for i in range(7):
        B.append((f"b{16+3*i}", prev, [_c(128, 1, 1, T, T, 16 + i * 3), _c(256, 3, 1, T, T, 17 + i * 3)], True)); prev = f"b{16+3*i}"
    B.append(("skip_36", prev, None, None))
    B.append(("b37", prev, [_c(512, 3, 2, T, T, 37), _c(256, 1, 1, T, T, 38), _c(512, 3, 1, T, T, 39)], True)); prev = "b37"
    for i in range(7):
        B.append((f"b{41+3*i}", prev, [_c(256, 1, 1, T, T, 41 + i * 3), _c(512, 3, 1, T, T, 42 + i * 3)], True)); prev = f"b{41+3*i}"
    B.append(("skip_61", prev, None, None))
    B.append(("b62", prev, [_c(1024, 3, 2, T, T, 62), _c(512, 1, 1, T, T, 63), _c(1024, 3, 1, T, T, 64)], True)); prev = "b62"
    for i in range(3):
        B.append((f"b{66+3*i}", prev, [_c(512, 1, 1, T, T, 66 + i * 3), _c(1024, 3, 1, T, T, 67 + i * 3)], True)); prev = f"b{66+3*i}"
    B.append(("base_out", prev, None, None))   # == add_46 / conv_73+add: FaceDetector base output
    B.append(("b75", prev, [_c(512, 1, 1, T, T, 75), _c(1024, 3, 1, T, T, 76), _c(512, 1, 1, T, T, 77),
                            _c(1024, 3, 1, T, T, 78), _c(512, 1, 1, T, T, 79)], False))
    B.append(("yolo_82", "b75", [_c(1024, 3, 1, T, T, 80), _c(head_c, 1, 1, Fa, Fa, 81)], False))
    B.append(("b84", "b75", [_c(256, 1, 1, T, T, 84)], False))
    B.append(("cat_a", ("up", "b84", "skip_61"), None, None))
    B.append(("b87", "cat_a", [_c(256, 1, 1, T, T, 87), _c(512, 3, 1, T, T, 88), _c(256, 1, 1, T, T, 89),
                               _c(512, 3, 1, T, T, 90), _c(256, 1, 1, T, T, 91)], False))
    B.append(("yolo_94", "b87", [_c(512, 3, 1, T, T, 92), _c(head_c, 1, 1, Fa, Fa, 93)], False))
    B.append(("b96", "b87", [_c(128, 1, 1, T, T, 96)], False))
    B.append(("cat_b", ("up", "b96", "skip_36"), None, None))
    B.append(("yolo_106", "cat_b", [_c(128, 1, 1, T, T, 99), _c(256, 3, 1, T, T, 100), _c(128, 1, 1, T, T, 101),
                                    _c(256, 3, 1, T, T, 102), _c(128, 1, 1, T, T, 103), _c(256, 3, 1, T, T, 104),
                                    _c(head_c, 1, 1, Fa, Fa, 105)], False))
    return B


def conv_list(head_c: int, fd6: bool = False, bb_info_c_size: int = 6):
    """[(layer_idx, cin, conv_dict)] in ascending layer index = weight-stream order (:91-119)."""
    out = []
    chans = {"in": 3}
    for name, src, convs, skip in graph_blocks(head_c):
        if convs is None:
            if isinstance(src, tuple):
                chans[name] = chans[src[1]] + chans[src[2]]
            else:
                chans[name] = chans[src]
            continue
        cin = chans[src]
        for cv in convs:
            out.append((cv["layer_idx"], cin, cv))
            cin = cv["filter"]
        chans[name] = cin
    out.sort(key=lambda t: t[0])
    if fd6:
        out = [t for t in out if t[0] <= 73]
        out.append((1000, 1024, {"filter": bb_info_c_size, "kernel": 3, "stride": 1, "bnorm": False, "leaky": False,
                                 "layer_idx": 1000}))   # face_detection.py:348-352
    return out


def parse_stream(stream: np.ndarray, head_c: int, fd6: bool = False, bb_info_c_size: int = 6):
    """Darknet float stream -> {idx: dict(kernel[kh,kw,cin,cout] (Keras layout), beta,gamma,mean,var | bias)}."""
    W = {}
    off = 0
    for idx, cin, cv in conv_list(head_c, fd6, bb_info_c_size):
        co, k = cv["filter"], cv["kernel"]
        d = {}
        if cv["bnorm"]:
            for nm in ("beta", "gamma", "mean", "var"):           # :97-101
                d[nm] = stream[off:off + co]; off += co
        else:
            d["bias"] = stream[off:off + co]; off += co            # :108-109
        n = co * cin * k * k
        kern = stream[off:off + n].reshape(co, cin, k, k); off += n    # :112, :117  (out,in,h,w)
        d["kernel_oihw"] = kern
        W[idx] = d
    if off != stream.size:
        raise ValueError(f"weight stream has {stream.size} floats, graph consumes {off}")
    return W


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _conv_block(x, convs, skip, W, emulate_bf16, taps, head_idxs):
    """yolov3_detect.py:196-215."""
    skip_connection = None
    for count, cv in enumerate(convs):
        if count == len(convs) - 2 and skip:
            skip_connection = x                                         # :201-202
        idx = cv["layer_idx"]
        w = torch.from_numpy(np.ascontiguousarray(W[idx]["kernel_oihw"]))
        last = count == len(convs) - 1
        if not emulate_bf16:
            if cv["kernel"] > 1:
                x = F.pad(x, (1, 1, 1, 1))                              # :205 ZeroPadding2D(1)
            bias = None if cv["bnorm"] else torch.from_numpy(np.ascontiguousarray(W[idx]["bias"]))
            x = F.conv2d(x, w, bias, stride=cv["stride"], padding=0)    # :206-211 'valid'
            if cv["bnorm"]:                                             # :212
                g, b, m, v = (torch.from_numpy(np.ascontiguousarray(W[idx][k])).view(1, -1, 1, 1) for k in ("gamma", "beta", "mean", "var"))
                x = g * (x - m) / torch.sqrt(v + BN_EPS) + b
            if cv["leaky"]:
                x = F.leaky_relu(x, 0.1)                                # :213
            if last and skip:
                x = skip_connection + x                                 # :215
        else:
            if cv["bnorm"]:
                g, b, m, v = (torch.from_numpy(np.ascontiguousarray(W[idx][k])) for k in ("gamma", "beta", "mean", "var"))
                sc = g / torch.sqrt(v + BN_EPS)
                w = w * sc.view(-1, 1, 1, 1)
                bias = b - m * sc
            else:
                bias = torch.from_numpy(np.ascontiguousarray(W[idx]["bias"]))
            w = _bf16(w)
            if cv["kernel"] > 1:
                x = F.pad(x, (1, 1, 1, 1))
            x = F.conv2d(x, w, bias, stride=cv["stride"], padding=0)
            if cv["leaky"]:
                x = F.leaky_relu(x, 0.1)
            if last and skip:
                x = skip_connection + x
            if idx not in head_idxs:
                x = _bf16(x)
        if taps is not None:
            taps[idx] = x
    return x


def forward(stream: np.ndarray, x_nhwc: np.ndarray, nb_class: int = 1, fd6: bool = False, bb_info_c_size: int = 6,
            emulate_bf16: bool = False, taps: dict | None = None, threads: int | None = None):
    """Run the restated graph.  Returns [yolo_82, yolo_94, yolo_106] (NHWC float32 numpy) or, for
    ``fd6``, the single (B,13,13,bb_info_c_size) map.  ``taps`` (dict) receives every conv's stored
    output (NCHW torch) keyed by layer index."""
    if threads:
        torch.set_num_threads(threads)
    head_c = 3 * (5 + nb_class)
    W = parse_stream(np.asarray(stream, np.float32), head_c, fd6, bb_info_c_size)
    head_idxs = (81, 93, 105, 1000)
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(x_nhwc, dtype=np.float32)).permute(0, 3, 1, 2).contiguous()
        if emulate_bf16:
            x = _bf16(x)
        vals = {"in": x}
        for name, src, convs, skip in graph_blocks(head_c):
            if fd6 and name == "b75":
                break
            if convs is None:
                if isinstance(src, tuple):
                    up = F.interpolate(vals[src[1]], scale_factor=2, mode="nearest")   # :282, :298 UpSampling2D(2)
                    vals[name] = torch.cat([up, vals[src[2]]], dim=1)                  # :283, :299
                else:
                    vals[name] = vals[src]
                continue
            vals[name] = _conv_block(vals[src], convs, skip, W, emulate_bf16, taps, head_idxs)
        if fd6:
            cv = {"filter": bb_info_c_size, "kernel": 3, "stride": 1, "bnorm": False, "leaky": False, "layer_idx": 1000}
            out = _conv_block(vals["base_out"], [cv], False, W, emulate_bf16, taps, head_idxs)   # 'same' 3x3 == pad 1 + valid
            return out.permute(0, 2, 3, 1).contiguous().numpy()
        return [vals[k].permute(0, 2, 3, 1).contiguous().numpy() for k in ("yolo_82", "yolo_94", "yolo_106")]
