"""Where does a network's gradient error come from?  Per-parameter rel-L2 error (in network order) of this library's
gradients and of the fp32 CPU oracle's, both against the fp64 oracle, on the 64x96 test network of tests/test_nets_gpu.py.

    python tools/diag_grad_error.py [disp|pose] [18|50] [fp32|tf32|tf32x3]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sc-sfmlearner-release_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch  # noqa: E402

import models  # noqa: E402
from golden_util import det_image, det_weights  # noqa: E402
from oracle import nets as N  # noqa: E402


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "disp"
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    mode = sys.argv[3] if len(sys.argv) > 3 else "fp32"
    img1, img2 = det_image("img1", 2, 64, 96), det_image("img2", 2, 64, 96)

    def loss_of(net, dt, dev):
        if kind == "disp":
            outs = net(img1.to(dt).to(dev))
            return sum(((1.0 / o) * (i + 1)).mean() for i, o in enumerate(outs)), outs
        o = net(img1.to(dt).to(dev), img2.to(dt).to(dev))
        return (o * torch.arange(1, 7, dtype=o.dtype, device=o.device)).sum() * 100, [o]

    def oracle(dt):
        ref = (N.DispResNet(layers) if kind == "disp" else N.PoseResNet(layers)).to(dt)
        ref.load_state_dict({k: v.to(dt) for k, v in det_weights(ref.state_dict()).items()})
        ref.train()
        acts = {}
        for name, m in ref.named_modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d)):
                m.register_forward_hook(lambda mod, i, o, name=name: acts.__setitem__(name, o.detach()))
        l, outs = loss_of(ref, dt, "cpu")
        l.backward()
        return {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}, [o.detach() for o in outs], acts
    g64, o64, a64 = oracle(torch.float64)
    g32, o32, a32 = oracle(torch.float32)
    net = models.DispResNet(layers, False) if kind == "disp" else models.PoseResNet(layers, False)
    net.load_state_dict(det_weights(net.state_dict()))
    net = net.cuda().set_conv_mode(mode).train()
    l, outs = loss_of(net, torch.float32, "cuda")
    l.backward()
    mine = {k: p.grad for k, p in net.named_parameters()}
    print("outputs: ours %s | fp32 oracle %s" % (["%.1e" % rel(a, b) for a, b in zip(outs, o64)], ["%.1e" % rel(a, b) for a, b in zip(o32, o64)]))
    print("forward activations of the fp32 oracle vs fp64 (conv / bn outputs), every 8th:")
    for i, k in enumerate(a64):
        if i % 8 == 0:
            print("   %-40s %.1e" % (k, rel(a32[k], a64[k])))
    print("%-55s %10s %10s %8s" % ("parameter", "ours", "fp32 cpu", "ratio"))
    for k in g64:
        e, r = rel(mine[k], g64[k]), rel(g32[k], g64[k])
        print("%-55s %10.2e %10.2e %8.1f%s" % (k, e, r, e / max(r, 1e-12), "   <<<" if e > 5 * r + 1e-5 else ""))


if __name__ == "__main__":
    main()
