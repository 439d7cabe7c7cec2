"""Per-role cycle counters of the persistent TMA convolution kernel (scsfm_conv_tma_debug) for a few layer shapes.
Usage: python tools/tma_profile.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sc-sfmlearner-release_b200"))
import torch

from scsfm import lib as L
from scsfm import nnops as O

O.CONFIG["conv_mode"] = "tf32"
lib = O._lib()
lib.scsfm_conv_tma_debug.argtypes = [ctypes.c_void_p]
NSM = torch.cuda.get_device_properties(0).multi_processor_count
CASES = [  # name, B, H, W, Cin, Cout, k, bn_groups, cfg
    ("enc L1 nobn", 12, 64, 208, 64, 64, 3, 0, (1, 2, 0, 0)),
    ("dec 0_1", 12, 256, 832, 16, 16, 3, 0, (1, 2, 0, 0)),
    ("enc L1", 12, 64, 208, 64, 64, 3, 3, (1, 2, 0, 0)),
    ("enc L2", 12, 32, 104, 128, 128, 3, 3, (1, 2, 0, 0)),
    ("dec 1_1", 12, 128, 416, 96, 32, 3, 0, (1, 2, 0, 0)),
]
_UNUSED = [
    ("enc L1", 12, 64, 208, 64, 64, 3, 3, (1, 1, 0, 0)),
    ("enc L1", 12, 64, 208, 64, 64, 3, 3, (1, 2, 0, 0)),
    ("enc L1 nobn", 12, 64, 208, 64, 64, 3, 0, (1, 1, 0, 0)),
    ("enc L2", 12, 32, 104, 128, 128, 3, 3, (1, 1, 0, 0)),
    ("enc L4", 12, 8, 26, 512, 512, 3, 3, (1, 1, 0, 0)),
    ("dec 0_1", 12, 256, 832, 16, 16, 3, 0, (1, 2, 0, 0)),
    ("dec 1_1", 12, 128, 416, 96, 32, 3, 0, (1, 2, 0, 0)),
    ("dec 2_1", 12, 64, 208, 128, 64, 3, 0, (1, 1, 0, 0)),
]
g = torch.Generator().manual_seed(0)
names = ["epi0 tmem-ld", "prod total", "mma wait-full", "mma wait-acc", "mma total", "epi wait-acc", "epi total", "tiles"]
for (name, B, H, W, Cin, Cout, k, groups, cfg) in CASES:
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).cuda()
    O.conv_tma_config(*cfg)
    sums = torch.zeros(O.BN_SLOTS * max(groups, 1) * Cout * 2, device="cuda", dtype=torch.float64) if groups else None
    for _ in range(2):
        O.conv_fwd(x, w, None, 1, k // 2, O.PAD_ZERO, O.ACT_NONE, sums, max(groups, 1))
    dbg = torch.zeros(NSM * 8, dtype=torch.int64, device="cuda")
    lib.scsfm_conv_tma_debug(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    O.conv_fwd(x, w, None, 1, k // 2, O.PAD_ZERO, O.ACT_NONE, sums, max(groups, 1))
    e1.record()
    torch.cuda.synchronize()
    lib.scsfm_conv_tma_debug(ctypes.c_void_p(0))
    d = dbg.view(NSM, 8).double()
    d = d[d[:, 7] > 0]
    m = d.mean(0)
    print("%-12s B%d %dx%d C%d->%d cfg%s  %.1f us, %d CTAs, %.1f tiles/CTA" % (name, B, H, W, Cin, Cout, cfg, e0.elapsed_time(e1) * 1e3, d.shape[0], m[7]))
    print("    " + "  ".join("%s=%.0f" % (n, v) for n, v in zip(names[:7], m[:7].tolist())), flush=True)
O.conv_tma_config(1)
