"""Per-layer times of the training-step convolution kernels against cuDNN (fp32 with TF32 off, and bf16).  usage: python tools/conv_bwd_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from face_vijnana_yolov3_b200 import conv_tc
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


B = 40
print("layer (Ci->Co k @HW, batch 40): microseconds   fwd: fvy | cudnn fp32 | cudnn bf16    dgrad: fvy | fp32 | bf16    wgrad: fvy | fp32 | bf16   GFLOP")
for ci, co, k, hw in [(32, 64, 3, 208), (64, 128, 3, 104), (128, 64, 1, 104), (128, 256, 3, 52), (256, 128, 1, 52), (256, 512, 3, 26), (512, 256, 1, 26),
                      (512, 1024, 3, 13), (1024, 512, 1, 13)]:
    x = torch.randn(B, ci, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, co, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
    w = (torch.randn(co, ci, k, k, device="cuda") * 0.05).contiguous()
    xb, dyb, wb = x.bfloat16(), dy.bfloat16(), w.bfloat16().contiguous(memory_format=torch.channels_last)
    p = k // 2
    r = []
    r.append(timeit(lambda: conv_tc.conv_forward(x, w)))
    r.append(timeit(lambda: F.conv2d(x, w, None, 1, p)))
    r.append(timeit(lambda: F.conv2d(xb, wb, None, 1, p)))
    r.append(timeit(lambda: conv_tc.conv_dgrad(dy, w)))
    r.append(timeit(lambda: torch.nn.grad.conv2d_input(x.shape, w, dy, 1, p)))
    r.append(timeit(lambda: torch.nn.grad.conv2d_input(xb.shape, wb, dyb, 1, p)))
    if ci % 64 == 0 and co % 64 == 0:
        r.append(timeit(lambda: conv_tc.conv_wgrad(x, dy, k)))
    else:
        r.append(float("nan"))
    r.append(timeit(lambda: torch.nn.grad.conv2d_weight(x, w.shape, dy, 1, p)))
    r.append(timeit(lambda: torch.nn.grad.conv2d_weight(xb, w.shape, dyb, 1, p)))
    gf = 2.0 * B * hw * hw * ci * co * k * k / 1e9
    print(f"{ci:4d}->{co:4d} k{k} @{hw:3d}:  fwd {r[0]:7.0f} | {r[1]:7.0f} | {r[2]:7.0f}    dgrad {r[3]:7.0f} | {r[4]:7.0f} | {r[5]:7.0f}    wgrad {r[6]:7.0f} | {r[7]:7.0f} | {r[8]:7.0f}   {gf:6.1f}", flush=True)
    conv_tc.clear_cache()
