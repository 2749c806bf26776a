import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from face_vijnana_yolov3_b200 import arch, synth, conv_tc, train as T
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
stream = synth.darknet_stream(arch.fd6_table(6), 0, synth.INIT_KERAS_DEFAULT)
hps = dict(lr=1e-4, beta_1=0.99, beta_2=0.99, decay=0.0)
x = torch.from_numpy(synth.images(2, 416, 416, 0)); t = torch.from_numpy(T.synthetic_targets(2, 1))
def run(mode, bn=None):
    tr = T.DataParallelTrainer(hps, device="cuda:0", stream=stream, fvy_conv_mode=mode, bucket_mb=8.0, fvy_bn=bn)
    loss = tr.step(x, t)
    g = [q.clone() for q in tr.flat_g]
    del tr
    return loss, g
def cmp(a, b):
    return [float((p.double() - q.double()).norm() / q.double().norm().clamp_min(1e-30)) for p, q in zip(a, b)]
r0 = run(0); r0b = run(0); r3 = run(3); r11 = run(11); r7 = run(7); r15 = run(15); rt = run(0, bn=False)
f = lambda v: " ".join("%.1e" % e for e in v)
print("losses", r0[0], r0b[0], r3[0], r11[0], r7[0], r15[0], rt[0])
print("fp32 rerun      ", f(cmp(r0b[1], r0[1])))
print("torch-bn vs fvy ", f(cmp(rt[1], r0[1])))
print("mode3 vs fp32   ", f(cmp(r3[1], r0[1])))
print("mode11 vs fp32  ", f(cmp(r11[1], r0[1])))
print("mode3 vs mode11 ", f(cmp(r3[1], r11[1])))
print("mode7 vs mode15 ", f(cmp(r7[1], r15[1])))
print("mode7 vs fp32   ", f(cmp(r7[1], r0[1])))
