"""Correctness + timing of the experimental weight-gradient kernels (scsfm_wgrad_config(1): wide cp.async kernel,
scsfm_wgrad_config(2): TMA kernel) against the default tcgen05 wgrad kernel and an fp64 torch reference, over the layer
shapes of DispResNet18 / PoseResNet18 at 256x832.

Usage: python tools/check_wgrad_wide.py [mode]      (mode 1 or 2, default 1)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sc-sfmlearner-release_b200"))
import torch
import torch.nn.functional as F

from scsfm import nnops as O

O.CONFIG["conv_mode"] = "tf32"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 1


def tf32(x):
    y = torch.empty_like(x)
    O.round_tf32(x.contiguous().view(-1), y.view(-1))
    return y


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# name, B, H, W, Cin, Cout, k, stride, pad, reflect
CASES = [
    ("enc L1", 12, 64, 208, 64, 64, 3, 1, 1, 0),
    ("enc L2", 12, 32, 104, 128, 128, 3, 1, 1, 0),
    ("enc L3", 12, 16, 52, 256, 256, 3, 1, 1, 0),
    ("enc L4", 12, 8, 26, 512, 512, 3, 1, 1, 0),
    ("enc L2 s2", 12, 64, 208, 64, 128, 3, 2, 1, 0),
    ("down 1x1 s2", 12, 64, 208, 64, 128, 1, 2, 0, 0),
    ("dec 0_1", 12, 256, 832, 16, 16, 3, 1, 1, 1),
    ("dec 0_0", 12, 128, 416, 32, 16, 3, 1, 1, 1),
    ("dec 1_1", 12, 128, 416, 96, 32, 3, 1, 1, 1),
    ("dec 2_1", 12, 64, 208, 128, 64, 3, 1, 1, 1),
    ("dec 4_1", 12, 16, 52, 512, 256, 3, 1, 1, 1),
    ("odd zero", 2, 37, 45, 20, 24, 3, 1, 1, 0),
    ("odd refl", 3, 19, 21, 36, 40, 3, 1, 1, 1),
    ("odd s2", 2, 37, 45, 24, 20, 3, 2, 1, 0),
    ("odd 1x1", 2, 9, 11, 68, 132, 1, 1, 0, 0),
]
g = torch.Generator().manual_seed(0)
bad = 0
for (name, B, H, W, Cin, Cout, k, s, pad, reflect) in CASES:
    x = tf32(torch.randn(B, H, W, Cin, generator=g).cuda())
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    dout = tf32(torch.randn(B, Ho, Wo, Cout, generator=g).cuda())
    pm = O.PAD_REFLECT if reflect else O.PAD_ZERO
    # fp64 reference through autograd
    xd = x.double().permute(0, 3, 1, 2)
    if reflect and pad:
        xd = F.pad(xd, (pad,) * 4, mode="reflect")
    wd = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(xd, wd, None, s, 0 if (reflect and pad) else pad).backward(dout.double().permute(0, 3, 1, 2))
    ref = wd.grad.permute(0, 2, 3, 1).contiguous()
    rb = dout.double().sum((0, 1, 2))

    def run(wide):
        O.wgrad_config(wide)
        dw = torch.zeros(Cout, k, k, Cin, device="cuda")
        db = torch.zeros(Cout, device="cuda")
        O.conv_wgrad(x, dout, dw, db, s, pad, pm)
        return dw, db

    dw0, db0 = run(0)
    t0 = timeit(lambda: run(0))
    dw1, db1 = run(MODE)
    torch.cuda.synchronize()
    t1 = timeit(lambda: run(MODE))
    e0, e1, d = rel(dw0, ref), rel(dw1, ref), rel(dw1, dw0)
    ok = e1 < 2e-3 and d < 1e-4 and rel(db1, rb) < 1e-4
    bad += 0 if ok else 1
    print("%-12s B%-2d %3dx%-3d C%3d->%-3d k%d s%d | err ref %.1e %.1e | mode-vs-default %.1e | %7.3f -> %7.3f ms  x%.2f %s"
          % (name, B, H, W, Cin, Cout, k, s, e0, e1, d, t0, t1, t0 / t1, "" if ok else "  <-- MISMATCH"), flush=True)
O.wgrad_config(0)
print("MISMATCHES: %d" % bad)
sys.exit(1 if bad else 0)
