import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sc-sfmlearner-release_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests/golden"))
import torch
from scsfm import nnops as O
O.CONFIG["conv_mode"] = "tf32"
g = torch.Generator().manual_seed(1)
for (B, H, W, Cin, Cout, k, s, pad) in [(2, 16, 24, 64, 64, 3, 1, 1), (2, 8, 12, 128, 128, 3, 1, 1), (2, 4, 6, 256, 256, 3, 1, 1), (2, 2, 3, 512, 512, 3, 1, 1), (2, 16, 24, 64, 128, 3, 2, 1), (2, 4, 6, 256, 512, 1, 2, 0)]:
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).cuda()
    y1 = O.conv_fwd(x, w, None, s, pad)
    x3 = torch.cat([x, x * 0.5, x + 1], 0).contiguous()
    y3 = O.conv_fwd(x3, w, None, s, pad)
    O.CONFIG["conv_mode"] = "fp32"
    yr = O.conv_fwd(x, w, None, s, pad)
    O.CONFIG["conv_mode"] = "tf32"
    print((B, H, W, Cin, Cout, k, s), "batched-vs-single max diff", float((y3[:B] - y1).abs().max()), " tf32-vs-fp32 rel", float((y1 - yr).norm() / yr.norm()),
          " mean signed rel bias", float(((y1 - yr) * yr.sign()).mean() / yr.abs().mean()))
# whole net: find first diverging activation between separate and batched
import models
from golden_util import det_image, det_weights
from scsfm import nets as NN
def build():
    n = models.DispResNet(18, False); n.load_state_dict(det_weights(n.state_dict())); return n.cuda().train()
a, b = build(), build()
imgs = [det_image(n, 2, 64, 96).cuda() for n in ("img1", "img2", "img3")]
a.ensure_arena(); b.ensure_arena()
with torch.no_grad():
    ra, _ = a._forward_impl(1, imgs[0])
    rb, _ = b._forward_impl(3, torch.cat(imgs, 0))
ea, eb = ra["enc"], rb["enc"]
def cmp(name, ta, tb):
    n = ta.shape[0]
    print("%-12s rel diff %.3e" % (name, float((tb[:n] - ta).norm() / ta.norm())))
cmp("y0", ea["y0"], eb["y0"]); cmp("f0", ea["f0"], eb["f0"])
for i, ((_, r1), (_, r2)) in enumerate(zip(ea["blocks"], eb["blocks"])):
    for key in ("y1", "h1", "y2", "out"):
        if key in r1:
            cmp("blk%d.%s" % (i, key), r1[key], r2[key])

for i in range(4, -1, -1):
    for key in ("a", "cat", "b"):
        cmp("dec%d.%s" % (i, key), ra["stages"][i][key], rb["stages"][i][key])
for k in ra["disps"]:
    cmp("disp%d" % k, ra["disps"][k], rb["disps"][k])
