"""Single convolutions on the tensor-core kernels (``fvy_conv_*``, include/fvy.h) for the training step.

``TcConv`` is one handle (one shape, one direction) of the tcgen05 implicit-GEMM kernel; ``conv_forward`` / ``conv_dgrad`` /
``conv_wgrad`` are what ``train._ConvFn`` calls: the forward of a 1 x 1 / 3 x 3 ``Conv2d`` (stride 1, or the backbone's stride-2
3 x 3), the gradient of a stride-1 ``Conv2d`` with respect to its input, dX = conv(dY, flip(W)^T), and the gradient with respect to
the weight - the work Keras / TensorFlow hand to cuDNN when the reference trains (face_detection.py:361-381, :602-630).  bf16
operands, fp32 accumulation and output; there is no CPU fallback.
"""
import ctypes as C
from typing import Dict, Tuple

import torch

from . import _lib as L


class TcConv:
    """cin -> cout, k x k (k = 1 or 3), zero padding k // 2, over (batch <= max_batch, height, width) INPUT maps; stride 1 or 2."""

    def __init__(self, device: int, height: int, width: int, cin: int, cout: int, k: int, max_batch: int, stride: int = 1):
        self.lib = L.load()
        self.shape = (height, width, cin, cout, k, stride)
        self.device, self.max_batch = device, max_batch
        h = C.c_void_p()
        L.check(self.lib.fvy_conv_create(device, height, width, cin, cout, k, stride, max_batch, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.fvy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, w: torch.Tensor, dgrad: bool) -> None:
        """w: the torch Conv2d weight (Co, Ci, k, k) float32 on this device, contiguous."""
        assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
        st = torch.cuda.current_stream(w.device).cuda_stream
        L.check(self.lib.fvy_conv_set_weights(self._h, C.c_void_p(w.data_ptr()), 1 if dgrad else 0, C.c_void_p(st)))

    def run(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, cin, H, W) float32 in channels_last memory format -> (B, cout, H / stride, W / stride) float32, channels_last."""
        height, width, cin, cout, _, stride = self.shape
        b = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (cin, height, width), (tuple(x.shape), self.shape)
        x = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty((b, cout, height // stride, width // stride), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        st = torch.cuda.current_stream(x.device).cuda_stream
        L.check(self.lib.fvy_conv_run(self._h, C.c_void_p(x.data_ptr()), b, C.c_void_p(y.data_ptr()), C.c_void_p(st)))
        return y


_cache: Dict[Tuple, TcConv] = {}
_scratch: Dict[Tuple, Tuple[int, torch.Tensor]] = {}


def _same_geometry(weight: torch.Tensor, padding: int) -> bool:
    co, ci, kh, kw = weight.shape
    return weight.is_cuda and weight.dtype == torch.float32 and kh == kw and kh in (1, 3) and padding == kh // 2


def eligible(weight: torch.Tensor, stride: int, padding: int) -> bool:
    """dgrad of this Conv2d can run on the tcgen05 kernel: stride 1, k in (1, 3) with 'same' padding, Co a multiple of 32
    (it is the K dimension of the dgrad GEMM), Ci <= 1024."""
    co, ci = weight.shape[:2]
    return _same_geometry(weight, padding) and stride == 1 and co % 32 == 0 and ci <= 1024 and ci % 4 == 0


def forward_eligible(weight: torch.Tensor, stride: int, padding: int) -> bool:
    co, ci, k = weight.shape[:3]
    if not _same_geometry(weight, padding) or co > 1024 or co % 4:
        return False
    return (stride == 1 and ci % 32 == 0) or (stride == 2 and k == 3 and ci % 64 == 0)


def wgrad_eligible(weight: torch.Tensor, stride: int, padding: int) -> bool:
    """Channel counts that are multiples of 64 run as they are; 32-channel operands (conv_2 / conv_3 of the backbone) are zero-padded
    to 64 channels on the way in (``conv_wgrad``).  Stride 2: the 3 x 3 layers with Co a multiple of 128."""
    co, ci, k = weight.shape[:3]
    if not _same_geometry(weight, padding):
        return False
    return (stride == 1 and ci % 32 == 0 and co % 32 == 0) or (stride == 2 and k == 3 and ci % 64 == 0 and co % 128 == 0)


def _handle(kind: str, dev: int, height: int, width: int, cin: int, cout: int, k: int, stride: int, batch: int) -> TcConv:
    key = (kind, dev, height, width, cin, cout, k, stride)
    h = _cache.get(key)
    if h is None or h.max_batch < batch:
        if h is not None:
            h.close()
        h = _cache[key] = TcConv(dev, height, width, cin, cout, k, max(batch, h.max_batch if h else 0), stride)
    return h


def _dev(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def conv_dgrad(dy: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Gradient w.r.t. the input of ``F.conv2d(x, weight, stride=1, padding=k // 2)`` given dy (B, Co, H, W)."""
    co, ci, k, _ = weight.shape
    b, _, height, width = dy.shape
    h = _handle("dgrad", _dev(dy), height, width, co, ci, k, 1, b)
    h.set_weights(weight.detach().contiguous(), dgrad=True)
    return h.run(dy)


def conv_forward(x: torch.Tensor, weight: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """``F.conv2d(x, weight, stride=stride, padding=k // 2)`` on the tcgen05 kernel (bf16 operands, fp32 accumulation and output)."""
    co, ci, k, _ = weight.shape
    b, _, height, width = x.shape
    h = _handle("fwd", _dev(x), height, width, ci, co, k, stride, b)
    h.set_weights(weight.detach().contiguous(), dgrad=False)
    return h.run(x)


def _scratch_for(dev: int, height: int, width: int, channels: int, batch: int, which: str, planes: int = 1) -> torch.Tensor:
    """Zeroed-once bf16 operand buffer of ``fvy_conv_wgrad`` for one (map, channels) shape; grown (and re-zeroed) for a larger batch.
    The four phase planes of a stride-2 layer's input are spaced by a batch-dependent row count: one buffer per batch size."""
    key = (dev, height, width, channels, which, planes, batch if planes > 1 else 0)
    have = _scratch.get(key)
    if have is None or have[0] < batch:
        rows = L.load().fvy_conv_wgrad_scratch_rows(batch, height, width)
        have = _scratch[key] = (batch, torch.zeros(planes * rows * channels, dtype=torch.bfloat16, device=f"cuda:{dev}"))
    return have[1]


def _pad_channels(t: torch.Tensor, c: int) -> torch.Tensor:
    out = torch.empty((t.shape[0], c, t.shape[2], t.shape[3]), dtype=t.dtype, device=t.device, memory_format=torch.channels_last).zero_()
    out[:, : t.shape[1]] = t
    return out


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, k: int, stride: int = 1) -> torch.Tensor:
    """Gradient w.r.t. the weight (Co, Ci, k, k) of ``F.conv2d(x, weight, stride=stride, padding=k // 2)`` given x (B, Ci, H, W) and
    dy (B, Co, H / stride, W / stride), both float32 (rounded once to bf16; fp32 accumulation)."""
    lib = L.load()
    ci_real, co_real = x.shape[1], dy.shape[1]
    if ci_real % 64:
        x = _pad_channels(x, (ci_real + 63) // 64 * 64)
    if co_real % 64:
        dy = _pad_channels(dy, (co_real + 63) // 64 * 64)
    b, ci = x.shape[:2]
    co, height, width = dy.shape[1:]
    dev = _dev(x)
    x = x.contiguous(memory_format=torch.channels_last)
    dy = dy.contiguous(memory_format=torch.channels_last)
    xs = _scratch_for(dev, height, width, ci, b, "x", 4 if stride == 2 else 1)
    ys = _scratch_for(dev, height, width, co, b, "dy")
    dw = torch.empty((co, ci, k, k), dtype=torch.float32, device=x.device)
    work = torch.empty_like(dw) if k > 1 else None
    st = torch.cuda.current_stream(x.device).cuda_stream
    P = lambda t: C.c_void_p(t.data_ptr())
    L.check(lib.fvy_conv_wgrad(P(x), P(dy), b, height, width, ci, co, k, stride, P(xs), P(ys), P(dw), P(work) if work is not None else None, C.c_void_p(st)))
    if (co, ci) != (co_real, ci_real):
        dw = dw[:co_real, :ci_real].contiguous()
    return dw


def clear_cache() -> None:
    for h in _cache.values():
        h.close()
    _cache.clear()
    _scratch.clear()
