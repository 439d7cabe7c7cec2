"""Oracle: differentiable geometry of the warp (restates reference inverse_warp.py).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  All functions are dtype-generic
(fp32 / fp64) and differentiable through torch autograd so that the same code
gives the gradient oracle.
"""
import torch


def rotation_from_euler(angles):
    """R = Rx(rx) @ Ry(ry) @ Rz(rz)  for angles [B,3] -> [B,3,3].

    Follows reference inverse_warp.py:77-112 (euler2mat): the three elementary
    rotations are built and multiplied in x.y.z order.
    """
    rx, ry, rz = angles[:, 0], angles[:, 1], angles[:, 2]
    zero = torch.zeros_like(rx)
    one = torch.ones_like(rx)
    cz, sz = torch.cos(rz), torch.sin(rz)
    cy, sy = torch.cos(ry), torch.sin(ry)
    cx, sx = torch.cos(rx), torch.sin(rx)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).view(-1, 3, 3)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).view(-1, 3, 3)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).view(-1, 3, 3)
    return Rx @ Ry @ Rz


def rotation_from_quat(q):
    """Unit quaternion (1, q) normalised -> rotation (reference inverse_warp.py:115-136)."""
    full = torch.cat([torch.ones_like(q[:, :1]), q], 1)
    full = full / full.norm(dim=1, keepdim=True)
    w, x, y, z = full.unbind(1)
    rows = [w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
            2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
            2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z]
    return torch.stack(rows, 1).view(-1, 3, 3)


def pose_to_matrix(vec, rotation_mode="euler"):
    """(tx,ty,tz,rx,ry,rz) [B,6] -> [R|t] [B,3,4]  (reference inverse_warp.py:139-154)."""
    t = vec[:, :3].unsqueeze(-1)
    if rotation_mode == "euler":
        R = rotation_from_euler(vec[:, 3:])
    elif rotation_mode == "quat":
        R = rotation_from_quat(vec[:, 3:])
    else:
        raise ValueError(rotation_mode)
    return torch.cat([R, t], 2)


def back_project(depth, K_inv):
    """cam[b,:,y,x] = K_inv[b] @ (x, y, 1)^T * depth[b,y,x]   (reference inverse_warp.py:29-44).

    depth [B,H,W], K_inv [B,3,3] -> [B,3,H,W].  Unlike the reference there is no
    module-global pixel grid cache (inverse_warp.py:5,39-40): the grid is rebuilt.
    """
    B, H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=depth.dtype, device=depth.device), torch.arange(W, dtype=depth.dtype, device=depth.device),
                            indexing="ij")
    pix = torch.stack([xs, ys, torch.ones_like(xs)], 0).view(1, 3, -1).expand(B, 3, -1)
    rays = (K_inv @ pix).view(B, 3, H, W)
    return rays * depth.unsqueeze(1)


def project(cam, rot, tr, padding_mode):
    """Camera points -> normalised sampling coords + clamped depth (reference inverse_warp.py:194-227).

    Returns xn, yn [B,H,W] and Z [B,1,H,W].  In 'zeros' mode any coordinate outside
    [-1,1] is overwritten with the constant 2 (which also cuts its gradient), exactly
    as the in-place masked assignment at inverse_warp.py:219-224 does.
    """
    B, _, H, W = cam.shape
    p = rot @ cam.reshape(B, 3, -1) + tr
    X, Y = p[:, 0], p[:, 1]
    Z = p[:, 2].clamp(min=1e-3)
    xn = 2 * (X / Z) / (W - 1) - 1
    yn = 2 * (Y / Z) / (H - 1) - 1
    if padding_mode == "zeros":
        two = torch.full_like(xn, 2.0)
        xn = torch.where(((xn > 1) | (xn < -1)).detach(), two, xn)
        yn = torch.where(((yn > 1) | (yn < -1)).detach(), two, yn)
    return xn.view(B, H, W), yn.view(B, H, W), Z.view(B, 1, H, W)


# When True, bilinear_sample / the SSIM box filter call the library kernels the reference itself calls
# (F.grid_sample, F.avg_pool2d) instead of the tap-by-tap restatement.  tests/test_oracle_golden.py checks the
# two agree; bench.py's CPU baseline uses the fast form so that the timed CPU path is the reference's own.
USE_LIBRARY_KERNELS = False


def bilinear_sample(src, xn, yn, padding_mode):
    """torch.nn.functional.grid_sample(src, (xn,yn), 'bilinear', padding_mode, align_corners=False)
    written out tap by tap (call sites: reference inverse_warp.py:262,267).

    Published definition (ATen GridSampler): pixel coordinate ix = ((xn+1)*W - 1)/2;
    'border' clips ix to [0, W-1] with zero gradient at/outside the limits; the four
    neighbours floor(ix), floor(ix)+1 (same in y) are blended with weights
    (1-fx)(1-fy) ...; any neighbour outside the image contributes value 0 (and no
    gradient).  src [B,C,H,W]; xn, yn [B,Ho,Wo] -> [B,C,Ho,Wo].
    """
    if USE_LIBRARY_KERNELS:
        return torch.nn.functional.grid_sample(src, torch.stack([xn, yn], -1), mode="bilinear", padding_mode=padding_mode,
                                               align_corners=False)
    B, C, H, W = src.shape
    ix = ((xn + 1) * W - 1) / 2
    iy = ((yn + 1) * H - 1) / 2
    if padding_mode == "border":
        inx = (ix > 0) & (ix < W - 1)
        iny = (iy > 0) & (iy < H - 1)
        ix = torch.where(inx, ix, ix.detach().clamp(0, W - 1))
        iy = torch.where(iny, iy, iy.detach().clamp(0, H - 1))
    elif padding_mode != "zeros":
        raise ValueError(padding_mode)
    x0 = torch.floor(ix.detach())
    y0 = torch.floor(iy.detach())
    fx = ix - x0
    fy = iy - y0
    flat = src.reshape(B, C, H * W)
    out = 0
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            xi = x0 + dx
            yi = y0 + dy
            ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long().view(B, 1, -1).expand(B, C, -1)
            val = torch.gather(flat, 2, idx).view(B, C, *xn.shape[1:])
            out = out + val * (wx * wy * ok.to(src.dtype)).unsqueeze(1)
    return out


def invert_intrinsics(K):
    """K^-1 (reference inverse_warp.py:253 uses torch.inverse)."""
    return torch.linalg.inv(K)


def inverse_warp2(img, depth, ref_depth, pose, intrinsics, padding_mode="zeros"):
    """Warp `img` / `ref_depth` (source view) into the target view given target depth and pose.

    Restates reference inverse_warp.py:230-269.  Returns (projected_img [B,3,H,W],
    valid_mask [B,1,H,W] float, projected_depth [B,1,H,W], computed_depth [B,1,H,W]).
    """
    cam = back_project(depth.squeeze(1), invert_intrinsics(intrinsics))
    proj = intrinsics @ pose_to_matrix(pose)
    xn, yn, Z = project(cam, proj[:, :, :3], proj[:, :, 3:], padding_mode)
    warped = bilinear_sample(img, xn, yn, padding_mode)
    valid = (torch.maximum(xn.abs(), yn.abs()) <= 1).unsqueeze(1).to(img.dtype)
    proj_depth = bilinear_sample(ref_depth, xn, yn, padding_mode)
    return warped, valid, proj_depth, Z


def inverse_warp(img, depth, pose, intrinsics, rotation_mode="euler", padding_mode="zeros"):
    """Legacy single-output warp (reference inverse_warp.py:157-191, cam2pixel :47-74):
    no out-of-range rewrite of the coordinates; returns (projected_img, valid bool [B,H,W])."""
    B, _, H, W = img.shape
    cam = back_project(depth, invert_intrinsics(intrinsics))
    proj = intrinsics @ pose_to_matrix(pose, rotation_mode)
    xn, yn, _ = project(cam, proj[:, :, :3], proj[:, :, 3:], "border")
    warped = bilinear_sample(img, xn, yn, padding_mode)
    valid = torch.maximum(xn.abs(), yn.abs()) <= 1
    return warped, valid
