"""Shared helpers for the parity tests."""
import numpy as np
import torch


def t(a, dtype=torch.float32, device="cpu"):
    return torch.from_numpy(np.asarray(a)).to(device=device, dtype=dtype)


def golden_loss_inputs(g, dtype=torch.float32, device="cpu", n_scales=2, requires_grad=False):
    """Rebuild the (tgt_img, ref_imgs, K, tgt_depth, ref_depths, poses, poses_inv) tuple of warp_loss.npz."""
    def leaf(a):
        x = t(a, dtype, device)
        return x.requires_grad_(True) if requires_grad else x
    tgt = t(g["in_tgt_img"], dtype, device)
    refs = [t(g[f"in_ref_img{i}"], dtype, device) for i in range(2)]
    K = t(g["in_K"], dtype, device)
    td = [leaf(g[f"in_tgt_depth_s{s}"]) for s in range(n_scales)]
    rd = [[leaf(g[f"in_ref_depth{i}_s{s}"]) for s in range(n_scales)] for i in range(2)]
    ps = [leaf(g[f"in_pose{i}"]) for i in range(2)]
    pi = [leaf(g[f"in_pose_inv{i}"]) for i in range(2)]
    return tgt, refs, K, td, rd, ps, pi


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def frac_within(a, b, tol):
    """Fraction of elements with |a-b| <= tol * max|b|."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float(((a - b).abs() <= tol * b.abs().max()).double().mean())
