import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from face_vijnana_yolov3_b200 import conv_tc, _lib as L
torch.manual_seed(0)
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
for (b, ci, co, Ho) in [(2, 64, 128, 8), (2, 64, 128, 20), (2, 64, 128, 52), (1, 64, 128, 104), (2, 64, 128, 104)]:
    Wo = Ho; H, W = 2 * Ho, 2 * Wo
    conv_tc.clear_cache()
    x = torch.randn(b, ci, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    dy = torch.randn(b, co, Ho, Wo, device="cuda").contiguous(memory_format=torch.channels_last)
    dw = conv_tc.conv_wgrad(x, dy, 3, stride=2)
    ref = torch.nn.grad.conv2d_weight(bf(x), (co, ci, 3, 3), bf(dy), 2, 1)
    total = L.load().fvy_conv_wgrad_scratch_rows(b, Ho, Wo)
    pitch = Wo + 1; plane = (Ho + 1) * pitch; lead = (pitch + 1 + 7) & ~7
    xs = [v[1] for k, v in conv_tc._scratch.items() if k[4] == "x"][0].float().reshape(4 * total, ci)
    ys = [v[1] for k, v in conv_tc._scratch.items() if k[4] == "dy"][0].float().reshape(total, co)
    X = torch.zeros((4 * total, ci), device="cuda"); Y = torch.zeros((total, co), device="cuda")
    n_i, y_i, x_i = torch.meshgrid(torch.arange(b), torch.arange(H), torch.arange(W), indexing="ij")
    hp, wp = y_i + 1, x_i + 1
    rows = lead + (((hp & 1) << 1) | (wp & 1)) * total + n_i * plane + (hp >> 1) * pitch + (wp >> 1)
    X[rows.flatten().cuda()] = bf(x).permute(0, 2, 3, 1).reshape(-1, ci)
    n_o, y_o, x_o = torch.meshgrid(torch.arange(b), torch.arange(Ho), torch.arange(Wo), indexing="ij")
    Y[(lead + n_o * plane + (y_o + 1) * pitch + (x_o + 1)).flatten().cuda()] = bf(dy).permute(0, 2, 3, 1).reshape(-1, co)
    bx = torch.nonzero((xs - X).abs().amax(1) > 0).flatten(); by = torch.nonzero((ys - Y).abs().amax(1) > 0).flatten()
    # emulate from the expected buffers
    rows_k = (b * plane + 63) // 64 * 64
    emu = torch.zeros(co, ci, 3, 3, device="cuda")
    for r in range(3):
        for q in range(3):
            off = (((r & 1) << 1) | (q & 1)) * total + ((r >> 1) - 1) * pitch + ((q >> 1) - 1)
            emu[:, :, r, q] = Y[lead:lead + rows_k].T @ X[lead + off:lead + off + rows_k]
    print(f"b={b} Ho={Ho} total={total} lead={lead}: err vs torch per tap", (dw - ref).abs().amax(dim=(0, 1)).flatten().cpu().numpy().round(3),
          "| emu vs torch", float((emu - ref).abs().max()), "| bad X rows", bx[:6].tolist(), len(bx), "bad Y rows", by[:6].tolist(), len(by), flush=True)
