"""Correctness + timing of the TMA halo-patch convolution kernel (csrc/conv_tma.cu) against the cp.async gather kernel
(csrc/conv_tc.cu) and an fp64 torch reference, over the layer shapes of DispResNet18 / PoseResNet18 at 256x832.

Usage: python tools/check_conv_tma.py [--quick]
Every line: case | config | rel-L2 error vs fp64 reference (gather, tma) | rel-L2 tma vs gather | time gather, tma (ms)
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
QUICK = "--quick" in sys.argv


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


def ref_fwd(x, w, bias, stride, pad, reflect, act):
    xd = x.double().permute(0, 3, 1, 2)
    wd = w.double().permute(0, 3, 1, 2)
    if reflect and pad:
        xd = F.pad(xd, (pad,) * 4, mode="reflect")
        pad = 0
    y = F.conv2d(xd, wd, bias.double() if bias is not None else None, stride, pad)
    if act == O.ACT_RELU:
        y = y.relu()
    elif act == O.ACT_ELU:
        y = F.elu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def ref_dgrad(dout, w, x_shape, stride, pad):
    B, Hi, Wi, Cin = x_shape
    dd = dout.double().permute(0, 3, 1, 2)
    wd = w.double().permute(0, 3, 1, 2)
    opad_h = Hi - ((dout.shape[1] - 1) * stride - 2 * pad + w.shape[1])
    opad_w = Wi - ((dout.shape[2] - 1) * stride - 2 * pad + w.shape[2])
    dx = F.conv_transpose2d(dd, wd, None, stride, pad, (opad_h, opad_w))
    return dx.permute(0, 2, 3, 1).contiguous()


CONFIGS = [("auto", (1, 0, 0, 0)), ("mt1", (1, 1, 0, 0)), ("mt2", (1, 2, 0, 0)), ("mt1 tw8", (1, 1, 0, 3)), ("mt1 tw16", (1, 1, 0, 4)),
           ("mt2 tw8", (1, 2, 0, 3)), ("mt2 tw16", (1, 2, 0, 4))]
if QUICK:
    CONFIGS = CONFIGS[:3]

# name, B, H, W, Cin, Cout, k, stride, pad, reflect, act, bias, bn_groups
FWD = [
    ("enc L1", 12, 64, 208, 64, 64, 3, 1, 1, 0, O.ACT_NONE, 0, 3),
    ("enc L2", 12, 32, 104, 128, 128, 3, 1, 1, 0, O.ACT_NONE, 0, 3),
    ("enc L3", 12, 16, 52, 256, 256, 3, 1, 1, 0, O.ACT_NONE, 0, 3),
    ("enc L4", 12, 8, 26, 512, 512, 3, 1, 1, 0, O.ACT_NONE, 0, 3),
    ("pose L1", 16, 64, 208, 64, 64, 3, 1, 1, 0, O.ACT_NONE, 0, 4),
    ("squeeze 1x1", 16, 8, 26, 512, 256, 1, 1, 0, 0, O.ACT_RELU, 1, 0),
    ("dec 0_1", 12, 256, 832, 16, 16, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 0_0", 12, 128, 416, 32, 16, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 1_1", 12, 128, 416, 96, 32, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 1_0", 12, 64, 208, 64, 32, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 2_1", 12, 64, 208, 128, 64, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 2_0", 12, 32, 104, 128, 64, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 3_1", 12, 32, 104, 256, 128, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 4_1", 12, 16, 52, 512, 256, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("dec 4_0", 12, 8, 26, 512, 256, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("odd zero", 2, 37, 45, 20, 24, 3, 1, 1, 0, O.ACT_RELU, 1, 2),
    ("odd refl", 3, 19, 21, 36, 40, 3, 1, 1, 1, O.ACT_ELU, 1, 0),
    ("odd 1x1", 2, 9, 11, 68, 132, 1, 1, 0, 0, O.ACT_NONE, 0, 0),
]
# name, B, Hi, Wi, Cin, Cout, k, stride, pad, padded_input
DGRAD = [
    ("enc L1", 12, 64, 208, 64, 64, 3, 1, 1, 0),
    ("enc L4", 12, 8, 26, 512, 512, 3, 1, 1, 0),
    ("dec 0_1 (padded)", 12, 256, 832, 16, 16, 3, 1, 1, 1),
    ("dec 1_1 (padded)", 12, 128, 416, 96, 32, 3, 1, 1, 1),
    ("dec 2_1 (padded)", 12, 64, 208, 128, 64, 3, 1, 1, 1),
    ("enc L2 s2", 12, 64, 208, 64, 128, 3, 2, 1, 0),
    ("down 1x1 s2", 12, 64, 208, 64, 128, 1, 2, 0, 0),
    ("odd s1", 2, 37, 45, 20, 24, 3, 1, 1, 0),
    ("odd s2", 2, 37, 45, 24, 20, 3, 2, 1, 0),
]
if QUICK:
    FWD = [FWD[0], FWD[3], FWD[6], FWD[8], FWD[15], FWD[16], FWD[17]]
    DGRAD = [DGRAD[0], DGRAD[3], DGRAD[5], DGRAD[7], DGRAD[8]]

g = torch.Generator().manual_seed(0)
bad = 0
print("== forward")
for (name, B, H, W, Cin, Cout, k, s, pad, reflect, act, has_bias, groups) in FWD:
    x = tf32(torch.randn(B, H, W, Cin, generator=g).cuda())
    w = tf32((torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).cuda())
    bias = torch.randn(Cout, generator=g).cuda() if has_bias else None
    pm = O.PAD_REFLECT if reflect else O.PAD_ZERO
    ref = ref_fwd(x, w, bias, s, pad, reflect, act)

    def run(with_sums=True):
        sums = torch.zeros(O.BN_SLOTS * groups * Cout * 2, device="cuda", dtype=torch.float64) if (groups and with_sums) else None
        y = O.conv_fwd(x, w, bias, s, pad, pm, act, sums, max(groups, 1))
        return y, sums

    O.conv_tma_config(0)
    y0, s0 = run()
    t0 = timeit(lambda: run(False))
    e0 = rel(y0, ref)
    for cname, cfg in CONFIGS:
        O.conv_tma_config(*cfg)
        y1, s1 = run()
        torch.cuda.synchronize()
        t1 = timeit(lambda: run(False))
        e1, d = rel(y1, ref), rel(y1, y0)
        ds = 0.0
        if s0 is not None:
            a = s0.view(O.BN_SLOTS, -1).sum(0)
            b_ = s1.view(O.BN_SLOTS, -1).sum(0)
            ds = rel(b_, a)
        ok = d < 2e-5 and ds < 1e-6 and e1 < 2e-3
        bad += 0 if ok else 1
        print("%-18s B%-2d %3dx%-3d C%3d->%-3d k%d %-8s | err ref %.1e %.1e | tma-vs-gather %.1e  bn %.1e | %7.3f -> %7.3f ms  x%.2f %s"
              % (name, B, H, W, Cin, Cout, k, cname, e0, e1, d, ds, t0, t1, t0 / t1, "" if ok else "  <-- MISMATCH"), flush=True)

print("== dgrad")
for (name, B, Hi, Wi, Cin, Cout, k, s, pad, padded) in DGRAD:
    w = tf32((torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).cuda())
    Ho, Wo = (Hi + 2 * pad - k) // s + 1, (Wi + 2 * pad - k) // s + 1
    dout = tf32(torch.randn(B, Ho, Wo, Cout, generator=g).cuda())
    addend = torch.randn(B, Hi, Wi, Cin, generator=g).cuda() if not padded else None
    if padded:
        ref = ref_dgrad(dout, w, (B, Hi + 2, Wi + 2, Cin), s, 0)
    else:
        ref = ref_dgrad(dout, w, (B, Hi, Wi, Cin), s, pad) + addend.double()

    O.invalidate_weight_cache()          # the flipped-weight cache is keyed by pointer: a freed `w` may be reused

    def run():
        return O.conv_dgrad(dout, w, (B, Hi, Wi, Cin), s, pad, addend, bool(padded))

    O.conv_tma_config(0)
    y0 = run()
    t0 = timeit(run)
    e0 = rel(y0, ref)
    for cname, cfg in CONFIGS:
        O.conv_tma_config(*cfg)
        y1 = run()
        torch.cuda.synchronize()
        t1 = timeit(run)
        e1, d = rel(y1, ref), rel(y1, y0)
        ok = d < 2e-5 and e1 < 2e-3
        bad += 0 if ok else 1
        print("%-18s B%-2d %3dx%-3d C%3d->%-3d k%d s%d %-8s | err ref %.1e %.1e | tma-vs-gather %.1e | %7.3f -> %7.3f ms  x%.2f %s"
              % (name, B, Hi, Wi, Cin, Cout, k, s, cname, e0, e1, d, t0, t1, t0 / t1, "" if ok else "  <-- MISMATCH"), flush=True)
O.conv_tma_config(1)
print("MISMATCHES: %d" % bad)
sys.exit(1 if bad else 0)
