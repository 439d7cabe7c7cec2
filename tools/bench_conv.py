"""Time individual conv layers (tensor-core kernels) with CUDA events; used under ncu for per-kernel analysis.
Usage: python tools/bench_conv.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sc-sfmlearner-release_b200"))
import torch
from scsfm import nnops as O
O.CONFIG["conv_mode"] = "tf32"
LAYERS = [  # B, H, W, Cin, Cout, k, stride, pad
    (12, 64, 208, 64, 64, 3, 1, 1),
    (12, 32, 104, 128, 128, 3, 1, 1),
    (12, 16, 52, 256, 256, 3, 1, 1),
    (12, 8, 26, 512, 512, 3, 1, 1),
    (12, 128, 416, 96, 32, 3, 1, 1),
    (12, 256, 832, 16, 16, 3, 1, 1),
]
g = torch.Generator().manual_seed(0)
for (B, H, W, Cin, Cout, k, s, pad) in LAYERS:
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).cuda()
    y = O.conv_fwd(x, w, None, s, pad)
    dy = torch.randn_like(y)
    dw = torch.zeros_like(w)
    res = {}
    for name, fn in (("fwd", lambda: O.conv_fwd(x, w, None, s, pad)), ("dgrad", lambda: O.conv_dgrad(dy, w, x.shape, s, pad)),
                     ("wgrad", lambda: O.conv_wgrad(x, dy, dw, None, s, pad))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * B * (H // s) * (W // s) * Cout * k * k * Cin
        res[name] = "%7.3f ms %6.1f TF/s" % (ms, fl / ms / 1e9)
    print("B%d %dx%d C%d->%d k%d s%d :" % (B, H, W, Cin, Cout, k, s), " | ".join("%s %s" % kv for kv in res.items()), flush=True)
