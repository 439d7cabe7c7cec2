"""Kernel-level timing of the fused loss kernels (CUDA events, L2 flushed between iterations).
Usage: python tools/bench_loss.py [B H W]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sc-sfmlearner-release_b200"))
import torch  # noqa: E402

from scsfm import synth  # noqa: E402
import loss_functions as lf  # noqa: E402


def main():
    B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 256, 832)
    d = synth.loss_inputs(0, B, H, W, n_ref=2, n_scales=1)
    c = lambda x: x.cuda()  # noqa: E731
    tgt, refs, K = c(d["tgt_img"]), [c(x) for x in d["ref_imgs"]], c(d["intrinsics"])
    td = [c(x).requires_grad_(True) for x in d["tgt_depth"]]
    rd = [[c(x).requires_grad_(True) for x in r] for r in d["ref_depths"]]
    ps = [c(x).requires_grad_(True) for x in d["poses"]]
    pi = [c(x).requires_grad_(True) for x in d["poses_inv"]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    res = {"pair_fwd": [], "pair_bwd": [], "smooth_fwd": [], "smooth_bwd": []}
    for it in range(13):
        flush.zero_()
        e = [ev() for _ in range(6)]
        e[0].record()
        p, q = lf.compute_photo_and_geometry_loss(tgt, refs, K, td, rd, ps, pi, 1, 1, 1, 1, "zeros")
        e[1].record()
        flush.zero_()
        e[2].record()
        s = lf.compute_smooth_loss(td, tgt, rd, refs)
        e[3].record()
        loss = p + 0.5 * q
        flush.zero_()
        e4, e5 = ev(), ev()
        e4.record()
        loss.backward()
        e5.record()
        flush.zero_()
        e6, e7 = ev(), ev()
        e6.record()
        s.backward()
        e7.record()
        torch.cuda.synchronize()
        if it >= 3:
            res["pair_fwd"].append(e[0].elapsed_time(e[1]))
            res["smooth_fwd"].append(e[2].elapsed_time(e[3]))
            res["pair_bwd"].append(e4.elapsed_time(e5))
            res["smooth_bwd"].append(e6.elapsed_time(e7))
        for t in td + [x for r in rd for x in r] + ps + pi:
            t.grad = None
    px = B * H * W
    alg = {"pair_fwd": 4 * 32 * px, "pair_bwd": 4 * 44 * px, "smooth_fwd": 3 * 16 * px, "smooth_bwd": 3 * 20 * px}
    for k, v in res.items():
        v.sort()
        med = v[len(v) // 2]
        print("%-11s median %.3f ms (min %.3f)  algorithmic %.1f MB -> %.0f GB/s (incl. launch/autograd overhead)"
              % (k, med, v[0], alg[k] / 1e6, alg[k] / med / 1e6))


if __name__ == "__main__":
    main()
