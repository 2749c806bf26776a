"""Training path (SURVEY 8e "Training", f-1): ground-truth tensors against the reference's own TrainingSequence output,
the torch restatement of the FaceDetector graph against the oracle forward, Keras Adam, and the world_size-2 gloo run of
the data-parallel step (gradient all-reduce)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200 import train as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gt_tensor_matches_reference_training_sequence(golden_dir):
    g = np.load(os.path.join(golden_dir, "gt_tensor.npz"))
    for i, (w, h) in enumerate(g["wh"]):
        faces = g["faces"][g["owner"] == i]
        faces = [f for f in faces if (f > 0).all()]
        assert np.array_equal(T.gt_tensor(faces, int(w), int(h)), g["targets"][i])
    # rows with a non-positive entry are skipped inside gt_tensor as well
    i = 0
    assert np.array_equal(T.gt_tensor(g["faces"][g["owner"] == i], int(g["wh"][i][0]), int(g["wh"][i][1])), g["targets"][i])


def test_training_sequence_reads_csv_like_the_reference(tmp_path):
    cv = pytest.importorskip("cv2")
    pd = pytest.importorskip("pandas")
    sizes = {"a.jpg": (320, 240), "b.jpg": (200, 260), "c.jpg": (128, 128)}
    for n, (w, h) in sizes.items():
        cv.imwrite(str(tmp_path / n), np.full((h, w, 3), 90, np.uint8))
    rows = [[0, "a.jpg", 1, 10, 20, 40, 50], [1, "a.jpg", 2, 100, 100, 0, 30], [2, "b.jpg", 3, 5, 5, 60, 80], [3, "c.jpg", 4, 30, 30, 20, 20]]
    pd.DataFrame(rows, columns=["FACE_ID", "FILE", "SUBJECT_ID", "FACE_X", "FACE_Y", "FACE_WIDTH", "FACE_HEIGHT"]).to_csv(tmp_path / "training.csv", index=False)
    hps = {"batch_size": 2}
    seq = T.TrainingSequence(str(tmp_path), hps, {"image_size": 416, "bb_info_c_size": 6})
    assert len(seq) == 2 and hps["step"] == 2
    x0, t0 = seq[0]
    x1, t1 = seq[1]
    assert x0.shape == (2, 416, 416, 3) and t0.shape == (2, 13, 13, 6) and x1.shape == (1, 416, 416, 3)
    assert x0.max() <= 1.0 and x0[0, 0, 0, 0] == 0.0            # letterbox border of the wide image
    assert int((t0[0, ..., 0] > 0).sum()) == 1                   # the zero-width face is skipped
    assert np.array_equal(t0[0], T.gt_tensor([rows[0][3:]], 320, 240))


def test_fdnet_matches_oracle_forward_and_stream_roundtrip():
    from oracle import darknet_ref as D
    specs = arch.fd6_table(6)
    stream = synth.darknet_stream(specs, 3, synth.INIT_BN_EXERCISING)
    x = synth.images(2, 64, 64, 1)
    net = T.FdNet(6)
    net.load_stream(stream)
    assert np.array_equal(net.to_stream(), stream)
    net.eval()
    with torch.no_grad():
        y = net(torch.from_numpy(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy()
    ref = D.forward(stream, x, 1, fd6=True)
    ref = ref[0] if isinstance(ref, (list, tuple)) else ref
    assert y.shape == ref.shape == (2, 2, 2, 6)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-4


def test_keras_adam_formula():
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(1000).astype(np.float32)
    grads = [rng.standard_normal(1000).astype(np.float32) for _ in range(4)]
    lr, b1, b2, decay, eps = 1e-3, 0.9, 0.999, 0.01, 1e-7
    # literal restatement of keras/optimizers.py (2.2.4) Adam.get_updates in float64
    p, m, v = p0.astype(np.float64), np.zeros(1000), np.zeros(1000)
    for it, g in enumerate(grads):
        lr_i = lr * (1.0 / (1.0 + decay * it))
        t = it + 1
        lr_t = lr_i * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g.astype(np.float64) ** 2
        p = p - lr_t * m / (np.sqrt(v) + eps)
    tp, tg = torch.from_numpy(p0.copy()), torch.zeros(1000)
    opt = T.FlatAdam([tp], [tg], lr, b1, b2, decay, eps)
    for g in grads:
        tg.copy_(torch.from_numpy(g) * 2.0)
        opt.step(grad_scale=0.5)
    assert np.allclose(tp.numpy(), p, rtol=2e-5, atol=2e-6)


def test_gloo_two_rank_training_step_matches_whole_batch_gradients(tmp_path):
    """World size 2 on CPU (gloo): parameters stay identical across ranks, and one step equals Adam applied to the MEAN of
    the two per-slice gradients (multi_gpu_model's whole-batch loss with per-tower BatchNorm statistics)."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys, numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        from face_vijnana_yolov3_b200 import arch, synth, train as T
        torch.set_num_threads(2)
        dist.init_process_group('gloo')
        r, n = dist.get_rank(), dist.get_world_size()
        hps = dict(lr=1e-3, beta_1=0.9, beta_2=0.999, decay=0.0)
        stream = synth.darknet_stream(arch.fd6_table(6), 0, synth.INIT_KERAS_DEFAULT)
        images = synth.images(4, 64, 64, 7)
        targets = np.random.default_rng(1).random((4, 2, 2, 6)).astype(np.float32)
        tr = T.DataParallelTrainer(hps, device='cpu', stream=stream, bucket_mb=8.0)
        assert len(tr.buckets) > 3
        xs, ts = T.slice_for_rank(images, targets, r, n)
        loss = tr.step(torch.from_numpy(xs), torch.from_numpy(ts))
        assert tr.last_allreduce_bytes == tr.n_params * 4
        # identical parameters on both ranks
        flat = torch.cat([p.reshape(-1) for p in tr.flat_p])
        both = [torch.empty_like(flat) for _ in range(n)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])
        if r == 0:
            # single-process restatement: per-slice gradients, averaged, one Keras-Adam step
            gs = []
            for k in range(n):
                net = T.FdNet(6); net.load_stream(stream); net.train()
                lo, hi = T.shard_bounds(4, n)[k]
                y = net(torch.from_numpy(images[lo:hi]).permute(0, 3, 1, 2))
                torch.nn.functional.mse_loss(y, torch.from_numpy(targets[lo:hi]).permute(0, 3, 1, 2)).backward()
                gs.append({{name: p.grad.clone() for name, p in net.named_parameters()}})
            net = T.FdNet(6); net.load_stream(stream)
            lr_t = T.keras_adam_lr_t(1e-3, 0.9, 0.999, 0.0, 0)
            worst = 0.0
            trained = dict(tr.model.named_parameters())
            for name, p in net.named_parameters():
                g = (gs[0][name] + gs[1][name]) / 2
                m = 0.1 * g; v = 0.001 * g * g
                want = p.detach() - lr_t * m / (v.sqrt() + 1e-7)
                worst = max(worst, float((trained[name].detach() - want).abs().max()))
            print('WORST', worst, 'LOSS', loss)
            assert worst < 1e-5, worst
        dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29653", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "WORST" in out.stdout


def test_gloo_unequal_and_empty_slices_weight_gradients_like_the_whole_batch(tmp_path):
    """A short last batch gives ranks unequal (3 images -> 2 + 1) or empty (1 image -> 1 + 0) slices.  multi_gpu_model takes the
    loss mean over the WHOLE batch (face_detection.py:366, :369): rank r's mean-loss gradient must enter with weight b_r / B
    (ADVICE r1), and a rank without images contributes zeros but still joins every all-reduce."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys, numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        from face_vijnana_yolov3_b200 import arch, synth, train as T
        torch.set_num_threads(2)
        dist.init_process_group('gloo')
        r, n = dist.get_rank(), dist.get_world_size()
        hps = dict(lr=1e-3, beta_1=0.9, beta_2=0.999, decay=0.0)
        stream = synth.darknet_stream(arch.fd6_table(6), 0, synth.INIT_KERAS_DEFAULT)
        for B in (3, 1):
            images = synth.images(B, 64, 64, 7)
            targets = np.random.default_rng(1).random((B, 2, 2, 6)).astype(np.float32)
            tr = T.DataParallelTrainer(hps, device='cpu', stream=stream, bucket_mb=8.0)
            xs, ts = T.slice_for_rank(images, targets, r, n)
            tr.step(torch.from_numpy(np.ascontiguousarray(xs)), torch.from_numpy(np.ascontiguousarray(ts)), global_batch=B)
            flat = torch.cat([p.reshape(-1) for p in tr.flat_p])
            both = [torch.empty_like(flat) for _ in range(n)]
            dist.all_gather(both, flat)
            assert torch.equal(both[0], both[1])
            if r == 0:
                acc = None
                for k in range(n):
                    lo, hi = T.shard_bounds(B, n)[k]
                    if hi == lo:
                        continue
                    net = T.FdNet(6); net.load_stream(stream); net.train()
                    y = net(torch.from_numpy(images[lo:hi]).permute(0, 3, 1, 2))
                    (torch.nn.functional.mse_loss(y, torch.from_numpy(targets[lo:hi]).permute(0, 3, 1, 2)) * ((hi - lo) / B)).backward()
                    g = {{name: p.grad.clone() for name, p in net.named_parameters()}}
                    acc = g if acc is None else {{k2: acc[k2] + g[k2] for k2 in g}}
                net = T.FdNet(6); net.load_stream(stream)
                lr_t = T.keras_adam_lr_t(1e-3, 0.9, 0.999, 0.0, 0)
                trained = dict(tr.model.named_parameters())
                worst = 0.0
                for name, p in net.named_parameters():
                    g = acc[name]
                    want = p.detach() - lr_t * (0.1 * g) / ((0.001 * g * g).sqrt() + 1e-7)
                    worst = max(worst, float((trained[name].detach() - want).abs().max()))
                print('WORST', B, worst)
                assert worst < 1e-5, (B, worst)
            dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29654", str(script)], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("WORST") == 2


@pytest.mark.gpu
def test_fused_adam_kernel_matches_torch_ops():
    p0 = torch.randn(100003, device="cuda")
    g = torch.randn(100003, device="cuda")
    a = T.FlatAdam([p0.clone()], [g.clone()], 1e-3, 0.9, 0.999, 0.01)
    b = T.FlatAdam([p0.clone()], [g.clone()], 1e-3, 0.9, 0.999, 0.01)
    b._lib = None                                  # torch-op restatement on the same device
    for _ in range(3):
        a.step(grad_scale=0.125); b.step(grad_scale=0.125)
    torch.cuda.synchronize()
    # torch's element-wise kernels contract a*b+c into FMAs, the library is built with -fmad=false: a few ulp apart
    assert torch.allclose(a.params[0], b.params[0], rtol=1e-5, atol=1e-6)
    assert torch.allclose(a.m[0], b.m[0], rtol=1e-5, atol=1e-7) and torch.allclose(a.v[0], b.v[0], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_trained_weights_flow_into_the_inference_engine():
    """Three optimizer steps on cuda:0, then the trained stream drives the tcgen05 engine: its head output matches the torch
    model in eval mode within the bf16 tolerance, and the loss went down."""
    from face_vijnana_yolov3_b200 import _lib as L
    from face_vijnana_yolov3_b200.engine import Engine
    hps = dict(lr=1e-4, beta_1=0.9, beta_2=0.999, decay=0.0)
    stream = synth.darknet_stream(arch.fd6_table(6), 0, synth.INIT_KERAS_DEFAULT)
    tr = T.DataParallelTrainer(hps, device="cuda:0", stream=stream)
    x = synth.images(4, 416, 416, 2)
    t = T.synthetic_targets(4, 3)
    losses = [tr.step(torch.from_numpy(x), torch.from_numpy(t)) for _ in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    new_stream = tr.weight_stream()
    assert not np.array_equal(new_stream, stream)
    eng = Engine(416, 416, head=L.HEAD_FD6, max_batch=4, bb_info_c_size=6)
    eng.load_weights(new_stream)
    out = eng.forward(x)
    out = out[0] if isinstance(out, (list, tuple)) else out
    tr.model.eval()
    with torch.no_grad():
        ref = tr.model(torch.from_numpy(x).cuda().permute(0, 3, 1, 2)).permute(0, 2, 3, 1).float().cpu().numpy()
    rel = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert rel < 2e-2, rel
    eng.close()


@pytest.mark.parametrize("stride,k", [(1, 3), (2, 3), (1, 1)])
def test_conv_fn_plain_path_equals_module_autograd(stride, k):
    """train._ConvFn with no kernel bit set (what every piece falls back to: explicit torch.nn.grad calls with the layer's stride and
    padding) gives the same output and gradients as nn.Conv2d's own autograd - CPU, float64-free, exact up to summation order."""
    import torch
    from face_vijnana_yolov3_b200 import train as T
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(8, 12, k, stride, padding=k // 2, bias=False)
    x1 = torch.randn(2, 8, 10, 10, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    w2 = conv.weight.detach().clone().requires_grad_(True)
    y1 = conv(x1)
    y2 = T._ConvFn.apply(x2, w2, k // 2, 0, stride)
    g = torch.randn_like(y1)
    y1.backward(g); y2.backward(g)
    assert torch.allclose(y1, y2, atol=1e-6)
    assert torch.allclose(x1.grad, x2.grad, atol=1e-5)
    assert torch.allclose(conv.weight.grad, w2.grad, atol=1e-4)


def test_conv_kernel_modes_are_inert_without_a_gpu():
    """On the CPU nothing is eligible for the tensor-core kernels: FdNet with every mode bit set takes the plain modules (no import of
    the CUDA library, same result as mode 0), and the eligibility helpers say no."""
    import torch
    from face_vijnana_yolov3_b200 import conv_tc, train as T
    torch.manual_seed(2)
    m = T.FdNet(6)
    m.train()
    x = torch.rand(1, 3, 64, 64)
    m.fvy_conv_mode = 0
    y0 = m(x)
    m.fvy_conv_mode = 7
    y7 = m(x)
    assert torch.equal(y0, y7)
    w = m.convs["3"].weight
    assert not conv_tc.eligible(w, 1, 1) and not conv_tc.wgrad_eligible(w, 1, 1) and not conv_tc.forward_eligible(w, 1, 1)
    assert not T._conv_fn_useful(m.convs["3"], 7)
