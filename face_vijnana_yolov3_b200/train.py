"""Data-parallel training step of the FaceDetector model (SURVEY section 8, rows (e) "Training" and f-1).

Replaces what the reference does with Keras in ``src/space/face_detection.py``:

    model  = Darknet-53 base conv_0..conv_73 + Conv2D(6, 3x3, 'same', linear)        :341-352, :384-600
    loss   = 'mse' against the (13, 13, 6) ground-truth tensor                       :366
    opt    = Adam(lr, beta_1, beta_2, decay)                                         :361-364
    multi_gpu_model(model, gpus=num_gpus) + fit_generator(TrainingSequence)          :330, :369, :602-630

B200 design: one process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch), the batch split into contiguous
slices exactly like ``multi_gpu_model``'s axis-0 split (``shard.shard_bounds``), weights and optimizer state replicated,
BatchNorm batch statistics per GPU (the reference's towers do the same; there is no SyncBN), and ONE exchange step: the
gradient all-reduce.  Gradients live in flat fp32 buckets (parameters' ``.grad`` are views into them); a bucket is
all-reduced asynchronously as soon as autograd has produced its last gradient, so the exchange overlaps the rest of the
backward pass; the optimizer waits for the handles.  Bucket size is chosen for launch latency, not link count (NVSwitch).

Status of row f-1: every BatchNormalization (training mode: batch statistics) + LeakyReLU pair runs through this repo's own
CUDA kernels forward and backward (``fvy_bn_leaky_train_forward / _backward``, csrc/train_kernels.cuh, bound to autograd by
``_BnLeakyFn``); the convolutions' forward / dgrad / wgrad still run through torch autograd (cuDNN - library code, the stated
baseline).  Also hand-written: the exchange (bucketing, overlap), the Keras-exact Adam update (``fvy_adam_step``, one fused
CUDA pass over the flat buckets) and the weight-stream interop that lets trained weights flow straight into the tcgen05
inference engine.  The ground-truth tensor builder restates ``TrainingSequence.__getitem__`` (:150-200).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import arch
from .shard import shard_bounds

KERAS_BN_MOMENTUM = 0.99     # Keras BatchNormalization default; torch's momentum is 1 - this
KERAS_EPSILON = 1e-7         # K.epsilon(), the Adam epsilon of Keras 2.2.4 when epsilon=None


# ----------------------------------------------------------------------------------------------------------------------
# ground truth (reference TrainingSequence.__getitem__, face_detection.py:150-200)
# ----------------------------------------------------------------------------------------------------------------------
def letterbox_geometry(w: int, h: int, image_size: int) -> Tuple[int, int, int, int]:
    """(w_p, h_p, pad_t, pad_l) of the reference's resize + copyMakeBorder (:113-140)."""
    if w >= h:
        w_p, h_p = image_size, int(h / w * image_size)
        return w_p, h_p, (image_size - h_p) // 2, 0
    h_p, w_p = image_size, int(w / h * image_size)
    return w_p, h_p, 0, (image_size - w_p) // 2


def gt_tensor(faces: Sequence[Sequence[float]], w: int, h: int, image_size: int = 416, cell_size: int = 13,
              bb_info_c_size: int = 6) -> np.ndarray:
    """Ground-truth (cell, cell, 6) tensor ``[obj, bx, by, bw, bh, cls]`` of one image.

    ``faces``: rows ``(FACE_X, FACE_Y, FACE_WIDTH, FACE_HEIGHT)`` in original-image pixels; rows with a non-positive
    entry are skipped (:147-149).  Later faces overwrite earlier ones that fall in the same cell, as in the reference.
    """
    cell_px = image_size // cell_size
    _, _, pad_t, pad_l = letterbox_geometry(w, h, image_size)
    gt = np.zeros((cell_size, cell_size, bb_info_c_size))
    for face in faces:
        if not all(v > 0 for v in face):
            continue
        x1, y1 = int(face[0]), int(face[1])
        x2, y2 = x1 + int(face[2]) - 1, y1 + int(face[3]) - 1
        wb, hb = x2 - x1 + 1, y2 - y1 + 1
        if w >= h:
            x1_p, y1_p = int(x1 / w * image_size), int(y1 / w * image_size) + pad_t
            x2_p, y2_p = int(x2 / w * image_size), int(y2 / w * image_size) + pad_t
        else:
            x1_p, y1_p = int(x1 / h * image_size) + pad_l, int(y1 / h * image_size)
            x2_p, y2_p = int(x2 / h * image_size) + pad_l, int(y2 / h * image_size)
        xc_p, yc_p = (x1_p + x2_p) // 2, (y1_p + y2_p) // 2
        cx, cy = xc_p // cell_px, yc_p // cell_px
        bx_p, by_p = (xc_p - cx * cell_px) / cell_px, (yc_p - cy * cell_px) / cell_px
        side = w if w >= h else h
        gt[cy, cx, :] = (1., bx_p, by_p, wb / side, hb / side, 1.)
    return gt


class TrainingSequence:
    """Batches of (letterboxed images, ground-truth tensors) from ``raw_data_path/training.csv`` - the reference's
    ``FaceDetector.TrainingSequence`` (face_detection.py:75-310): same file order (pandas groupby on FILE), same last
    short batch, cubic resize + zero border, images scaled to [0, 1]."""

    def __init__(self, raw_data_path, hps, nn_arch, cell_size=13):
        import os
        import pandas as pd
        self.raw_data_path, self.hps, self.nn_arch, self.cell_size = raw_data_path, hps, nn_arch, cell_size
        self.gt_df_g = pd.read_csv(os.path.join(raw_data_path, 'training.csv')).groupby('FILE')
        self.file_names = list(self.gt_df_g.groups.keys())
        self.batch_size = hps['batch_size']
        hps['step'] = -(-len(self.file_names) // self.batch_size)            # :87-90

    def __len__(self):
        return self.hps['step']

    def __getitem__(self, index):
        import os
        import cv2 as cv
        S = self.nn_arch['image_size']
        images, gts = [], []
        for bi in range(index * self.batch_size, min((index + 1) * self.batch_size, len(self.file_names))):
            name = self.file_names[bi]
            df = self.gt_df_g.get_group(name)
            image = cv.imread(os.path.join(self.raw_data_path, name), cv.IMREAD_COLOR)[:, :, ::-1] / 255
            h, w = image.shape[0], image.shape[1]
            w_p, h_p, pad_t, pad_l = letterbox_geometry(w, h, S)
            image = cv.resize(image, (w_p, h_p), interpolation=cv.INTER_CUBIC)
            image = cv.copyMakeBorder(image, pad_t, S - h_p - pad_t, pad_l, S - w_p - pad_l, cv.BORDER_CONSTANT, value=[0, 0, 0])
            faces = df[['FACE_X', 'FACE_Y', 'FACE_WIDTH', 'FACE_HEIGHT']].to_numpy()
            valid = (df.iloc[:, 3:] > 0).all(axis=1).to_numpy()                  # :147-149 looks at every column from the 4th on
            images.append(image)
            gts.append(gt_tensor([f for f, ok in zip(faces, valid) if ok], w, h, S, self.cell_size, self.nn_arch['bb_info_c_size']))
        return np.asarray(images), np.asarray(gts)


def synthetic_targets(batch: int, seed: int = 0, cell_size: int = 13, positives: float = 0.01) -> np.ndarray:
    """Synthetic (B, 13, 13, 6) targets with ~1 % positive cells."""
    rng = np.random.default_rng(seed)
    t = np.zeros((batch, cell_size, cell_size, 6), np.float32)
    pos = rng.random((batch, cell_size, cell_size)) < positives
    n = int(pos.sum())
    t[pos] = np.concatenate([np.ones((n, 1)), rng.random((n, 2)), 0.05 + 0.3 * rng.random((n, 2)), np.ones((n, 1))], 1)
    return t


# ----------------------------------------------------------------------------------------------------------------------
# BatchNormalization (training mode) + LeakyReLU on this repo's kernels
# ----------------------------------------------------------------------------------------------------------------------
class _BnLeakyFn(torch.autograd.Function):
    """y = LeakyReLU_slope(BatchNorm_train(x)) through ``fvy_bn_leaky_train_forward / _backward`` (NHWC fp32, CUDA only).
    The running statistics of ``bn`` (a torch BatchNorm2d used as the parameter / buffer holder) are updated in place."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, slope):
        from . import _lib as L
        lib = L.load()
        x = x.contiguous(memory_format=torch.channels_last)
        n, c, h, w = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        mean = torch.empty(c, dtype=torch.float32, device=x.device)
        invstd = torch.empty_like(mean)
        ws = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        P = lambda t: C.c_void_p(t.data_ptr())
        L.check(lib.fvy_bn_leaky_train_forward(P(x), C.c_longlong(n * h * w), c, P(gamma), P(beta), C.c_float(bn.eps), C.c_float(bn.momentum),
                                               C.c_float(slope), P(bn.running_mean), P(bn.running_var), P(y), P(mean), P(invstd), P(ws), C.c_void_p(st)))
        ctx.save_for_backward(x, gamma, beta, mean, invstd)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import _lib as L
        lib = L.load()
        x, gamma, beta, mean, invstd = ctx.saved_tensors
        dy = dy.contiguous(memory_format=torch.channels_last)
        n, c, h, w = x.shape
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        dgamma = torch.empty_like(gamma); dbeta = torch.empty_like(beta)
        ws = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        P = lambda t: C.c_void_p(t.data_ptr())
        L.check(lib.fvy_bn_leaky_train_backward(P(x), P(dy), C.c_longlong(n * h * w), c, P(gamma), P(beta), P(mean), P(invstd), C.c_float(ctx.slope),
                                                P(dx), P(dgamma), P(dbeta), P(ws), C.c_void_p(st)))
        return dx, dgamma, dbeta, None, None


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _tc_eligible(conv: nn.Conv2d) -> bool:
    from . import conv_tc
    return conv_tc.eligible(conv.weight, conv.stride[0], conv.padding[0])


def _conv_fn_useful(conv: nn.Conv2d, mode: int) -> bool:
    """At least one of the pieces `mode` asks for can run on this repo's kernels for this layer (otherwise the plain module is used:
    one fused cuDNN backward instead of two separate calls)."""
    from . import conv_tc
    w, s, p = conv.weight, conv.stride[0], conv.padding[0]
    return bool(((mode & 1) and conv_tc.eligible(w, s, p)) or ((mode & 2) and conv_tc.wgrad_eligible(w, s, p)) or
                ((mode & 4) and conv_tc.forward_eligible(w, s, p)))


class _ConvFn(torch.autograd.Function):
    """A stride-1 Conv2d on this repo's kernels, bf16 operands with fp32 accumulation (``conv_tc``).  ``mode`` bits: 1 = the gradient
    w.r.t. the INPUT on the tcgen05 implicit-GEMM kernel (``conv_dgrad``), 2 = the gradient w.r.t. the WEIGHT on ``conv_wgrad_kernel``
    (tcgen05 with MN-major operands; warp-level MMAs for 64-wide layers), 4 = the forward on the tcgen05 kernel as well.  The
    backbone's stride-2 3 x 3 layers take bits 2 and 4 (their dgrad stays on cuDNN).  Whatever a bit leaves out runs on cuDNN in fp32.
    Bit 8 (tests): the same arithmetic emulated with torch - fp32 library convolutions on operands rounded to bf16 - in place of the
    kernels, so that a whole step on the kernels can be checked against a reference that differs only in summation order."""

    @staticmethod
    def forward(ctx, x, weight, padding, mode, stride=1):
        from . import conv_tc
        ctx.save_for_backward(x, weight)
        ctx.padding, ctx.mode, ctx.stride = padding, mode, stride
        if (mode & 4) and conv_tc.forward_eligible(weight, stride, padding):
            if mode & 8:
                return F.conv2d(_bf16_round(x), _bf16_round(weight), None, stride, padding)
            return conv_tc.conv_forward(x, weight, stride)
        return F.conv2d(x, weight, None, stride, padding)

    @staticmethod
    def backward(ctx, dy):
        from . import conv_tc
        x, weight = ctx.saved_tensors
        dx = dw = None
        emu = bool(ctx.mode & 8)
        stride = ctx.stride
        if ctx.needs_input_grad[0]:
            if (ctx.mode & 1) and conv_tc.eligible(weight, stride, ctx.padding):
                dx = (torch.nn.grad.conv2d_input(x.shape, _bf16_round(weight), _bf16_round(dy), 1, ctx.padding) if emu
                      else conv_tc.conv_dgrad(dy, weight))
            else:
                dx = torch.nn.grad.conv2d_input(x.shape, weight, dy, stride, ctx.padding)
        if ctx.needs_input_grad[1]:
            if (ctx.mode & 2) and conv_tc.wgrad_eligible(weight, stride, ctx.padding):
                dw = (torch.nn.grad.conv2d_weight(_bf16_round(x), weight.shape, _bf16_round(dy), stride, ctx.padding) if emu
                      else conv_tc.conv_wgrad(x, dy, weight.shape[2], stride))
            else:
                dw = torch.nn.grad.conv2d_weight(x, weight.shape, dy, stride, ctx.padding)
        return dx, dw, None, None, None


# ----------------------------------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------------------------------
class FdNet(nn.Module):
    """The FaceDetector graph as a torch module, driven by the same layer table as the CUDA plan (``arch.fd6_table``).

    Keras semantics kept: ZeroPadding2D(1) + 'valid' conv for 3x3 (symmetric pad 1, also for stride 2), BatchNormalization
    (eps 1e-3, momentum 0.99) in training mode, LeakyReLU(0.1), residual add after the activation, linear 3x3 'same' head.
    """

    def __init__(self, bb_info_c_size: int = 6):
        super().__init__()
        self.specs = arch.fd6_table(bb_info_c_size)
        self.fvy_bn = False        # True: BatchNorm (training) + LeakyReLU through this repo's kernels (fp32 CUDA tensors only)
        self.fvy_dgrad = False     # True: input gradients of the stride-1 convolutions on the tcgen05 kernel (bf16 operands)
        self.fvy_conv_mode = 0     # _ConvFn mode bits (1 dgrad, 2 wgrad, 4 forward on this repo's kernels); fvy_dgrad = True is mode 1
        self.convs = nn.ModuleDict()
        self.bns = nn.ModuleDict()
        for c in self.specs:
            self.convs[str(c.idx)] = nn.Conv2d(c.cin, c.cout, c.k, c.stride, padding=1 if c.k == 3 else 0, bias=not c.bn)
            if c.bn:
                self.bns[str(c.idx)] = nn.BatchNorm2d(c.cout, eps=1e-3, momentum=1.0 - KERAS_BN_MOMENTUM)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, 3, S, S) in [0, 1] -> (B, 6, S/32, S/32)."""
        outs: Dict[int, torch.Tensor] = {-1: x}
        y = x
        for c in self.specs:
            conv = self.convs[str(c.idx)]
            xin = outs[c.src]
            mode = self.fvy_conv_mode | (1 if self.fvy_dgrad else 0)
            if (mode and self.training and conv.bias is None and xin.is_cuda and xin.dtype == torch.float32 and c.src != -1 and
                    _conv_fn_useful(conv, mode)):
                y = _ConvFn.apply(xin, conv.weight, conv.padding[0], mode, conv.stride[0])
            else:
                y = conv(xin)
            if c.bn and self.fvy_bn and self.training and y.is_cuda and y.dtype == torch.float32:
                bn = self.bns[str(c.idx)]
                y = _BnLeakyFn.apply(y, bn.weight, bn.bias, bn, 0.1 if c.leaky else 1.0)
            else:
                if c.bn:
                    y = self.bns[str(c.idx)](y)
                if c.leaky:
                    y = F.leaky_relu(y, 0.1)
            if c.res is not None:
                y = y + outs[c.res]
            outs[c.idx] = y
        return y

    # ---- Darknet stream interop (same order as WeightReader.load_weights, yolov3_detect.py:91-119)
    def load_stream(self, stream: np.ndarray) -> None:
        s = torch.from_numpy(np.ascontiguousarray(stream, np.float32))
        if s.numel() != arch.n_params(self.specs):
            raise ValueError(f"weight stream has {s.numel()} floats, model needs {arch.n_params(self.specs)}")
        off = 0

        def take(n):
            nonlocal off
            v = s[off:off + n]
            off += n
            return v

        with torch.no_grad():
            for c in self.specs:
                conv = self.convs[str(c.idx)]
                if c.bn:
                    bn = self.bns[str(c.idx)]
                    bn.bias.copy_(take(c.cout)); bn.weight.copy_(take(c.cout))
                    bn.running_mean.copy_(take(c.cout)); bn.running_var.copy_(take(c.cout))
                else:
                    conv.bias.copy_(take(c.cout))
                conv.weight.copy_(take(c.n_kernel).view_as(conv.weight))

    def to_stream(self) -> np.ndarray:
        parts = []
        for c in self.specs:
            conv = self.convs[str(c.idx)]
            if c.bn:
                bn = self.bns[str(c.idx)]
                parts += [bn.bias, bn.weight, bn.running_mean, bn.running_var]
            else:
                parts.append(conv.bias)
            parts.append(conv.weight.reshape(-1))
        return torch.cat([p.detach().float().reshape(-1).cpu() for p in parts]).numpy()


# ----------------------------------------------------------------------------------------------------------------------
# Keras Adam (keras/optimizers.py of 2.2.4): lr_t = lr / (1 + decay * iterations) * sqrt(1 - b2^t) / (1 - b1^t);
# m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t * m / (sqrt(v) + eps)
# ----------------------------------------------------------------------------------------------------------------------
def keras_adam_lr_t(lr: float, beta_1: float, beta_2: float, decay: float, iterations: int) -> float:
    t = iterations + 1
    lr_i = lr * (1.0 / (1.0 + decay * iterations)) if decay > 0 else lr
    return lr_i * math.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t)


class FlatAdam:
    """Keras-exact Adam over flat fp32 buckets.  On a GPU the update is ONE fused CUDA pass per bucket (``fvy_adam_step`` of
    libfvy.so, called through the C ABI with device pointers); on CPU (gloo tests) the same arithmetic in torch ops."""

    def __init__(self, params: List[torch.Tensor], grads: List[torch.Tensor], lr, beta_1, beta_2, decay, epsilon=KERAS_EPSILON):
        self.params, self.grads = params, grads
        self.lr, self.b1, self.b2, self.decay, self.eps = float(lr), float(beta_1), float(beta_2), float(decay), float(epsilon)
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.iterations = 0
        self._lib = None
        if params and params[0].is_cuda:
            from . import _lib as L
            self._lib = L.load()                     # raises if the CUDA library is missing: no silent fallback on a GPU box

    def step(self, grad_scale: float = 1.0) -> None:
        lr_t = keras_adam_lr_t(self.lr, self.b1, self.b2, self.decay, self.iterations)
        for p, g, m, v in zip(self.params, self.grads, self.m, self.v):
            if self._lib is not None:
                from . import _lib as L
                stream = torch.cuda.current_stream(p.device).cuda_stream
                L.check(self._lib.fvy_adam_step(C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(m.data_ptr()),
                                                C.c_void_p(v.data_ptr()), C.c_longlong(p.numel()), C.c_float(lr_t), C.c_float(self.b1),
                                                C.c_float(self.b2), C.c_float(self.eps), C.c_float(grad_scale), C.c_void_p(stream)))
            else:
                gs = g * grad_scale if grad_scale != 1.0 else g
                # (1 - beta) is a float32 subtraction in the Keras graph (beta is a float32 variable) and in the CUDA kernel
                c1 = float(np.float32(1.0) - np.float32(self.b1)); c2 = float(np.float32(1.0) - np.float32(self.b2))
                m.mul_(self.b1).add_(gs, alpha=c1)
                v.mul_(self.b2).addcmul_(gs, gs, value=c2)
                p.sub_(lr_t * m / (v.sqrt() + self.eps))
        self.iterations += 1


# ----------------------------------------------------------------------------------------------------------------------
# data-parallel trainer
# ----------------------------------------------------------------------------------------------------------------------
class DataParallelTrainer:
    """One rank of the data-parallel FaceDetector training job.

    ``hps``: the reference's ``fd_conf['hps']`` (lr, beta_1, beta_2, decay).  ``bucket_mb``: flat gradient bucket size.
    World size 1 (or no initialised process group) trains locally with the same code path minus the all-reduce.
    """

    def __init__(self, hps: dict, device: str = "cpu", bb_info_c_size: int = 6, bucket_mb: float = 32.0, autocast_bf16: bool = False,
                 stream: Optional[np.ndarray] = None, model: Optional[nn.Module] = None, fvy_bn: Optional[bool] = None,
                 fvy_dgrad: bool = False, fvy_conv_mode: int = 0):
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.model = model if model is not None else FdNet(bb_info_c_size)
        if stream is not None:
            self.model.load_stream(stream)
        self.model.to(self.device)
        if self.device.type == "cuda":
            self.model.to(memory_format=torch.channels_last)
        self.model.train()
        self.autocast_bf16 = autocast_bf16 and self.device.type == "cuda"
        # BatchNorm + LeakyReLU on this repo's kernels whenever they can run (fp32 on a GPU); torch's modules otherwise
        if hasattr(self.model, "fvy_bn"):
            self.model.fvy_bn = (self.device.type == "cuda" and not self.autocast_bf16) if fvy_bn is None else bool(fvy_bn)
        # input gradients of the stride-1 convolutions on the tcgen05 kernel (bf16 operands: opt-in, the reference trains in fp32)
        if hasattr(self.model, "fvy_dgrad"):
            self.model.fvy_dgrad = bool(fvy_dgrad) and self.device.type == "cuda" and not self.autocast_bf16
            self.model.fvy_conv_mode = int(fvy_conv_mode) if (self.device.type == "cuda" and not self.autocast_bf16) else 0
        # flat buckets in REVERSE parameter order (= the order autograd finishes gradients in)
        params = [p for p in self.model.parameters() if p.requires_grad]
        limit = int(bucket_mb * (1 << 20) / 4)
        self.buckets: List[List[torch.Tensor]] = [[]]
        n = 0
        for p in reversed(params):
            if n and n + p.numel() > limit:
                self.buckets.append([]); n = 0
            self.buckets[-1].append(p); n += p.numel()
        self.flat_p, self.flat_g = [], []
        self._bucket_of: Dict[torch.Tensor, int] = {}
        for bi, bucket in enumerate(self.buckets):
            total = sum(p.numel() for p in bucket)
            fp = torch.empty(total, dtype=torch.float32, device=self.device)
            fg = torch.zeros(total, dtype=torch.float32, device=self.device)
            off = 0
            for p in bucket:
                fp[off:off + p.numel()].copy_(p.data.reshape(-1))
                p.data = fp[off:off + p.numel()].view(p.shape)      # parameters and gradients are views into the flat buffers
                p.grad = fg[off:off + p.numel()].view(p.shape)
                self._bucket_of[p] = bi
                off += p.numel()
            self.flat_p.append(fp); self.flat_g.append(fg)
        self.n_params = sum(f.numel() for f in self.flat_p)
        self.opt = FlatAdam(self.flat_p, self.flat_g, hps['lr'], hps['beta_1'], hps['beta_2'], hps.get('decay', 0.0))
        self._pending = [0] * len(self.buckets)
        self._next_bucket = 0
        self._handles: List = []
        self._comm_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        if self.world > 1:
            for p in params:
                p.register_post_accumulate_grad_hook(self._on_grad)
        self.last_allreduce_bytes = 0
        self.exchange = True       # False: skip the all-reduce (bench.py measures the exposed communication as the difference)

    # -- exchange: a bucket goes out as soon as its last gradient has been accumulated
    def _on_grad(self, p: torch.Tensor) -> None:
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        self._launch_ready()

    def _launch_ready(self, force: bool = False) -> None:
        """Buckets go out strictly in index order (bucket i only after buckets < i), whatever order autograd completes them in:
        every rank - including one that runs no backward because its slice is empty - issues the same sequence of collectives."""
        while self._next_bucket < len(self.buckets) and (force or self._pending[self._next_bucket] == 0):
            self._launch(self._next_bucket)
            self._next_bucket += 1

    def exchange_all(self) -> None:
        """The step's exchange on its own (every bucket, same order and stream): bench.py times it for the bus bandwidth."""
        self._handles = []
        self._next_bucket = 0
        self._launch_ready(force=True)
        for h in self._handles:
            h.wait()
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)

    def _launch(self, bi: int) -> None:
        if not self.exchange:
            return
        fg = self.flat_g[bi]
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._comm_stream):
                h = dist.all_reduce(fg, op=dist.ReduceOp.SUM, async_op=True)
        else:
            h = dist.all_reduce(fg, op=dist.ReduceOp.SUM, async_op=True)
        self._handles.append(h)
        self.last_allreduce_bytes += fg.numel() * 4

    def step(self, images: torch.Tensor, targets: torch.Tensor, global_batch: Optional[int] = None) -> float:
        """One optimizer step on this rank's slice.  images (b, S, S, 3) NHWC float in [0,1]; targets (b, 13, 13, 6).
        Returns this rank's loss (Keras 'mse': mean over every element of the slice).

        ``multi_gpu_model`` takes the loss mean over the WHOLE concatenated batch (face_detection.py:366, :369), so a rank whose
        slice holds b of the B = ``global_batch`` images contributes its mean-loss gradient with weight b / B; with equal slices
        (the default when ``global_batch`` is None) that is the plain average over ranks.  A rank with an empty slice (short last
        batch) contributes zeros but still joins every all-reduce."""
        b_local = int(images.shape[0])
        weight = 1.0 if global_batch is None else b_local * self.world / float(global_batch)
        if b_local == 0:
            for g in self.flat_g:
                g.zero_()
            self._handles, self.last_allreduce_bytes = [], 0
            if self.world > 1:
                self.exchange_all()
            self.opt.step(grad_scale=1.0 / self.world)
            return 0.0
        x = images.to(self.device, non_blocking=True).permute(0, 3, 1, 2).float()
        if self.device.type == "cuda":
            x = x.contiguous(memory_format=torch.channels_last)
        t = targets.to(self.device, non_blocking=True).permute(0, 3, 1, 2).float()
        for g in self.flat_g:
            g.zero_()
        self._pending = [len(b) for b in self.buckets]
        self._handles, self.last_allreduce_bytes, self._next_bucket = [], 0, 0
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast_bf16):
            y = self.model(x)
        loss = F.mse_loss(y.float(), t)
        (loss * weight if weight != 1.0 else loss).backward()
        if self.world > 1:
            self._launch_ready(force=True)       # nothing is left in practice; keeps the collective sequence complete
        for h in self._handles:
            h.wait()
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        # multi_gpu_model: the loss is the mean over the WHOLE batch -> sum over ranks of (b_r / B) x the per-slice mean gradient
        self.opt.step(grad_scale=1.0 / self.world)
        return float(loss.detach())

    def weight_stream(self) -> np.ndarray:
        """Darknet-order stream of the current weights (BatchNorm running statistics averaged over ranks), ready for
        ``FaceDetector.set_weight_stream`` / ``Engine.load_weights``."""
        if self.world > 1:
            for bn in self.model.bns.values():
                for buf in (bn.running_mean, bn.running_var):
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                    buf.div_(self.world)
        return self.model.to_stream()


def slice_for_rank(images: np.ndarray, targets: np.ndarray, rank: int, world: int):
    """multi_gpu_model's axis-0 split: rank r gets the contiguous slice ``shard_bounds(B, world)[r]``."""
    lo, hi = shard_bounds(images.shape[0], world)[rank]
    return images[lo:hi], targets[lo:hi]
