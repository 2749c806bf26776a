"""Training-step timing (BASELINE.json configs[3]: FaceDetector training step, batch 40 @416, data-parallel with the NCCL gradient
all-reduce).  Launch: python tools/train_bench.py [--steps 10] or, for N GPUs,
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py --global-batch 40
Prints one JSON line on rank 0: step time (max over ranks, CUDA events), images/s, all-reduce bytes per step, and the
stand-alone all-reduce time / bus bandwidth of the same buckets (2 (N-1)/N x bytes / time).  Synthetic images and targets."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_vijnana_yolov3_b200 import arch, synth, train as T   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=40)
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bucket-mb", type=float, default=32.0)
    ap.add_argument("--fp32", action="store_true", help="no bf16 autocast (the reference trains in fp32)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.backends.cudnn.benchmark = True
    hps = dict(lr=1e-4, beta_1=0.99, beta_2=0.99, decay=0.0)
    stream = synth.darknet_stream(arch.fd6_table(6), 0, synth.INIT_KERAS_DEFAULT)
    tr = T.DataParallelTrainer(hps, device=f"cuda:{local}", stream=stream, bucket_mb=a.bucket_mb, autocast_bf16=not a.fp32)
    images = synth.images(a.global_batch, a.size, a.size, 0)
    targets = T.synthetic_targets(a.global_batch, 1, cell_size=a.size // 32)
    xs, ts = T.slice_for_rank(images, targets, rank, world)
    xs, ts = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ts).pin_memory()
    for _ in range(a.warmup):
        loss = tr.step(xs, ts)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = tr.step(xs, ts)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    # stand-alone exchange of the same buckets
    ar_ms = torch.zeros(1, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        for _ in range(2):
            for g in tr.flat_g:
                dist.all_reduce(g)
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for _ in range(5):
            for g in tr.flat_g:
                dist.all_reduce(g)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
        dist.all_reduce(ar_ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        nbytes = tr.n_params * 4
        out = {"metric": "FaceDetector training step (fwd+bwd+allreduce+Adam)", "n_gpus": world, "global_batch": a.global_batch, "net": a.size,
               "ms_per_step": float(ms), "images_per_s": a.global_batch / float(ms) * 1e3, "loss": loss,
               "dtype": "fp32" if a.fp32 else "bf16 autocast, fp32 master weights / gradients / Adam",
               "params": tr.n_params, "allreduce_bytes_per_step": nbytes if world > 1 else 0, "buckets": len(tr.buckets),
               "allreduce_alone_ms": float(ar_ms) if world > 1 else None,
               "allreduce_bus_GBps": (2 * (world - 1) / world * nbytes / (float(ar_ms) * 1e-3) / 1e9) if world > 1 else None,
               "compute": "torch autograd (cuDNN) - row f-1 of SURVEY 8 replaces it piecewise; exchange, Adam kernel and weight interop are this repo's",
               "data": "synthetic"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
