"""One eager training step of a bench configuration between cudaProfilerStart / cudaProfilerStop (for ncu / compute-sanitizer):

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/launches.csv python tools/one_step.py [tf32x3|tf32|fp32] [kitti_r18|kitti_r50|nyu_r18] [warmup]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sc-sfmlearner-release_b200"))
import torch  # noqa: E402

import models  # noqa: E402
from scsfm import synth  # noqa: E402
from scsfm.trainer import Trainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
dl, pl, H, W, n_ref, B, kind = {"kitti_r18": (18, 18, 256, 832, 2, 4, "kitti"), "kitti_r50": (50, 50, 256, 832, 2, 2, "kitti"),
                                "nyu_r18": (18, 18, 256, 320, 1, 8, "nyu"), "small": (18, 18, 128, 160, 2, 2, "kitti")}[sys.argv[2] if len(sys.argv) > 2 else "kitti_r18"]
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = "cuda"
torch.manual_seed(0)
tr = Trainer(models.DispResNet(dl, False).to(dev).train(), models.PoseResNet(pl, False).to(dev).train(), with_auto_mask=1,
             distributed=False, conv_mode=mode)
tgt, refs, K = synth.triplet(0, B, H, W, n_ref, kind)
args = (tgt.to(dev), [r.to(dev) for r in refs], K.to(dev))
for _ in range(warm):
    tr.step(*args)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = tr.step(*args)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one step [%s]: loss %.5f photo %.5f smooth %.5f geo %.5f" % ((mode,) + tuple(float(v) for v in out)))
