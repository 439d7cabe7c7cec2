"""Per-launch CUDA-event profile of one training step: every conv call with its shape, time, TFLOP/s.
Usage: python tools/profile_layers.py [fp32|tf32|tf32x3] [kitti_r18|kitti_r50|nyu_r18]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sc-sfmlearner-release_b200"))
import torch  # noqa: E402

import models  # noqa: E402
from scsfm import lib as L, synth  # noqa: E402
from scsfm.trainer import Trainer  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
    dl, pl, H, W, n_ref, B, kind = {"kitti_r18": (18, 18, 256, 832, 2, 4, "kitti"), "kitti_r50": (50, 50, 256, 832, 2, 2, "kitti"),
                                    "nyu_r18": (18, 18, 256, 320, 1, 8, "nyu")}[sys.argv[2] if len(sys.argv) > 2 else "kitti_r18"]
    dev = "cuda"
    tr = Trainer(models.DispResNet(dl, False).to(dev).train(), models.PoseResNet(pl, False).to(dev).train(),
                 with_auto_mask=1, distributed=False, conv_mode=mode)
    tgt, refs, K = synth.triplet(0, B, H, W, n_ref, kind)
    args = (tgt.to(dev), [r.to(dev) for r in refs], K.to(dev))
    for _ in range(2):
        tr.step(*args)
    torch.cuda.synchronize()
    L.PROF.update(enabled=True, only=None, events=[])
    tr.step(*args)
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for fam, work, e0, e1, tag in L.PROF["events"]:
        key = (fam, tag)
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += work
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    total = sum(v[1] for v in agg.values())
    fams = collections.OrderedDict()
    for (fam, _), (n, ms, _) in rows:
        f = fams.setdefault(fam, [0, 0.0])
        f[0] += n; f[1] += ms
    print("total profiled %.2f ms  [%s]" % (total, mode))
    print("  ".join("%s %.2f" % (k, v[1]) for k, v in sorted(fams.items(), key=lambda kv: -kv[1][1])))
    for (fam, tag), (n, ms, work) in rows[:70]:
        rate = work / (ms * 1e-3) / 1e12 if fam.startswith("conv") else work / (ms * 1e-3) / 1e9
        unit = "TF/s" if fam.startswith("conv") else "GB/s"
        print("%-16s %-40s n=%3d %8.3f ms %8.1f %s" % (fam, tag or "", n, ms, rate, unit))


if __name__ == "__main__":
    main()
