"""Oracle: photometric / geometry-consistency / smoothness losses (restates reference
loss_functions.py).  TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import torch
import torch.nn.functional as F

from . import geometry
from .geometry import inverse_warp2

SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2
MIN_MASK_SUM = 10000


def _box3(x):
    """3x3 mean over a reflect-padded-by-1 image ([B,C,H,W] -> [B,C,H,W]).

    ReflectionPad2d(1) followed by AvgPool2d(3, 1) (reference loss_functions.py:17-23):
    padded index -1 mirrors to 1 and H mirrors to H-2 (edge sample not repeated).
    """
    p = F.pad(x, (1, 1, 1, 1), mode="reflect")
    if geometry.USE_LIBRARY_KERNELS:
        return F.avg_pool2d(p, 3, 1)
    H, W = x.shape[-2:]
    acc = 0
    for dy in range(3):
        for dx in range(3):
            acc = acc + p[..., dy:dy + H, dx:dx + W]
    return acc / 9


def ssim_dissimilarity(x, y):
    """clamp((1 - SSIM(x,y)) / 2, 0, 1) with 3x3 box statistics (reference loss_functions.py:25-42)."""
    mu_x, mu_y = _box3(x), _box3(y)
    var_x = _box3(x * x) - mu_x * mu_x
    var_y = _box3(y * y) - mu_y * mu_y
    cov = _box3(x * y) - mu_x * mu_y
    num = (2 * mu_x * mu_y + SSIM_C1) * (2 * cov + SSIM_C2)
    den = (mu_x * mu_x + mu_y * mu_y + SSIM_C1) * (var_x + var_y + SSIM_C2)
    return ((1 - num / den) / 2).clamp(0, 1)


def mean_on_mask(diff, valid_mask):
    """sum(diff*mask)/sum(mask) over the whole batch if the expanded mask sum exceeds 10000,
    else the constant 0 (reference loss_functions.py:123-129)."""
    mask = valid_mask.expand_as(diff)
    if mask.sum() > MIN_MASK_SUM:
        return (diff * mask).sum() / mask.sum()
    return torch.zeros((), dtype=diff.dtype, device=diff.device)


def pairwise_terms(tgt_img, ref_img, tgt_depth, ref_depth, pose, intrinsic,
                   with_ssim, with_mask, with_auto_mask, padding_mode):
    """Per-pixel maps of one pair-direction (reference loss_functions.py:95-113).

    Returns dict with warped, valid (after auto-mask), proj_depth, comp_depth,
    diff_img (final photometric map [B,3,H,W]) and diff_depth [B,1,H,W].
    """
    warped, valid, proj_depth, comp_depth = inverse_warp2(ref_img, tgt_depth, ref_depth, pose,
                                                          intrinsic, padding_mode)
    diff_img = (tgt_img - warped).abs().clamp(0, 1)
    diff_depth = ((comp_depth - proj_depth).abs() / (comp_depth + proj_depth)).clamp(0, 1)
    warp_valid = valid
    if with_auto_mask == True:  # noqa: E712  (reference compares the int flag with == True)
        still = (tgt_img - ref_img).abs().mean(1, keepdim=True)
        valid = (diff_img.mean(1, keepdim=True) < still).to(valid.dtype) * valid
    if with_ssim == True:  # noqa: E712
        diff_img = 0.15 * diff_img + 0.85 * ssim_dissimilarity(tgt_img, warped)
    if with_mask == True:  # noqa: E712
        diff_img = diff_img * (1 - diff_depth)
    return dict(warped=warped, warp_valid=warp_valid, valid=valid, proj_depth=proj_depth,
                comp_depth=comp_depth, diff_img=diff_img, diff_depth=diff_depth)


def compute_pairwise_loss(tgt_img, ref_img, tgt_depth, ref_depth, pose, intrinsic,
                          with_ssim, with_mask, with_auto_mask, padding_mode):
    """(photometric, geometry) masked means of one pair-direction (reference loss_functions.py:95-119)."""
    t = pairwise_terms(tgt_img, ref_img, tgt_depth, ref_depth, pose, intrinsic,
                       with_ssim, with_mask, with_auto_mask, padding_mode)
    return mean_on_mask(t["diff_img"], t["valid"]), mean_on_mask(t["diff_depth"], t["valid"])


def _nearest_up(x, size):
    """F.interpolate(x, size, mode='nearest') for an exact power-of-two factor: src = dst >> s
    (reference loss_functions.py:81-82)."""
    return F.interpolate(x, size, mode="nearest")


def compute_photo_and_geometry_loss(tgt_img, ref_imgs, intrinsics, tgt_depth, ref_depths, poses, poses_inv,
                                    max_scales, with_ssim, with_mask, with_auto_mask, padding_mode):
    """Sum over refs x scales x both directions (reference loss_functions.py:50-92)."""
    photo = 0
    geo = 0
    H, W = tgt_img.shape[-2:]
    n_scales = min(len(tgt_depth), max_scales)
    for ref_img, ref_depth, pose, pose_inv in zip(ref_imgs, ref_depths, poses, poses_inv):
        for s in range(n_scales):
            td = tgt_depth[s] if s == 0 else _nearest_up(tgt_depth[s], (H, W))
            rd = ref_depth[s] if s == 0 else _nearest_up(ref_depth[s], (H, W))
            p1, g1 = compute_pairwise_loss(tgt_img, ref_img, td, rd, pose, intrinsics,
                                           with_ssim, with_mask, with_auto_mask, padding_mode)
            p2, g2 = compute_pairwise_loss(ref_img, tgt_img, rd, td, pose_inv, intrinsics,
                                           with_ssim, with_mask, with_auto_mask, padding_mode)
            photo = photo + (p1 + p2)
            geo = geo + (g1 + g2)
    return photo, geo


def smooth_term(depth, img):
    """Edge-aware first-order smoothness of the mean-normalised map (reference loss_functions.py:133-152).

    depth [B,1,H,W] is divided by (its per-image mean + 1e-7); |d/dx| and |d/dy| are weighted by
    exp(-mean_c |dI|) and averaged separately (x over B*H*(W-1), y over B*(H-1)*W).
    """
    d = depth / (depth.mean(2, True).mean(3, True) + 1e-7)
    gx = (d[..., :, :-1] - d[..., :, 1:]).abs()
    gy = (d[..., :-1, :] - d[..., 1:, :]).abs()
    wx = torch.exp(-(img[..., :, :-1] - img[..., :, 1:]).abs().mean(1, keepdim=True))
    wy = torch.exp(-(img[..., :-1, :] - img[..., 1:, :]).abs().mean(1, keepdim=True))
    return (gx * wx).mean() + (gy * wy).mean()


def compute_smooth_loss(tgt_depth, tgt_img, ref_depths, ref_imgs):
    """Scale-0 smoothness of the target and every reference depth (reference loss_functions.py:132-159)."""
    loss = smooth_term(tgt_depth[0], tgt_img)
    for rd, ri in zip(ref_depths, ref_imgs):
        loss = loss + smooth_term(rd[0], ri)
    return loss


@torch.no_grad()
def compute_errors(gt, pred, dataset):
    """Validation depth metrics with Garg/NYU crop and median scaling (reference loss_functions.py:163-205)."""
    B, H, W = gt.shape
    if dataset == "kitti":
        ys, xs, cap = (0.40810811, 0.99189189), (0.03594771, 0.96405229), 80
    elif dataset == "nyu":
        ys, xs, cap = (0.09375, 0.98125), (0.0640625, 0.9390625), 10
    else:
        raise ValueError(dataset)
    crop = torch.zeros(H, W, dtype=torch.bool)
    crop[int(ys[0] * H):int(ys[1] * H), int(xs[0] * W):int(xs[1] * W)] = True
    tot = torch.zeros(6, dtype=torch.float64)
    for g, p in zip(gt, pred):
        sel = (g > 0.1) & (g < cap) & crop
        vg = g[sel]
        vp = p[sel].clamp(1e-3, cap)
        vp = vp * torch.median(vg) / torch.median(vp)
        ratio = torch.max(vg / vp, vp / vg)
        tot += torch.stack([(vg - vp).abs().mean(), ((vg - vp).abs() / vg).mean(), ((vg - vp) ** 2 / vg).mean(),
                            (ratio < 1.25).float().mean(), (ratio < 1.25 ** 2).float().mean(),
                            (ratio < 1.25 ** 3).float().mean()]).double()
    return [float(v) / B for v in tot]
