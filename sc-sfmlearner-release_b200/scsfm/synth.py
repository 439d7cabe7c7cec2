"""Seeded synthetic KITTI-/NYU-shaped inputs.

Pure torch on CPU so that the oracle, the parity tests and bench.py all see the
same numbers.  Images are smoothed uniform noise normalised like the reference's
dataset transform (train.py:92-93: mean 0.45, std 0.225); disparities span the
decoder's output range (DispResNet.py:98) and poses the pose head's scale
(PoseResNet.py:49).
"""
import torch
import torch.nn.functional as F

KITTI_K = (483.4, 492.6, 408.3, 118.0, 256, 832)   # fx, fy, cx, cy at H, W
NYU_K = (259.4, 277.0, 162.8, 135.3, 256, 320)


def intrinsics(B, H, W, kind="kitti"):
    fx, fy, cx, cy, h0, w0 = KITTI_K if kind == "kitti" else NYU_K
    sx, sy = W / w0, H / h0
    K = torch.tensor([[fx * sx, 0.0, cx * sx], [0.0, fy * sy, cy * sy], [0.0, 0.0, 1.0]])
    return K.unsqueeze(0).repeat(B, 1, 1).contiguous()


def image(gen, B, H, W):
    u = torch.rand(B, 3, H + 8, W + 8, generator=gen)
    return ((F.avg_pool2d(u, 9, 1) - 0.45) / 0.225).contiguous()


def depth_map(gen, B, H, W):
    n = torch.randn(B, 1, H + 14, W + 14, generator=gen)
    disp = 10 * torch.sigmoid(3 * F.avg_pool2d(n, 15, 1)) + 0.01
    return (1 / disp).contiguous()


def pose(gen, B):
    return (0.01 * torch.randn(B, 6, generator=gen)).contiguous()


def triplet(seed, B, H, W, n_ref=2, kind="kitti"):
    """One batch in the dataset's return convention (datasets/sequence_folders.py:55-65):
    tgt_img, [ref_imgs], intrinsics."""
    g = torch.Generator().manual_seed(seed)
    tgt = image(g, B, H, W)
    refs = [image(g, B, H, W) for _ in range(n_ref)]
    return tgt, refs, intrinsics(B, H, W, kind)


def loss_inputs(seed, B, H, W, n_ref=2, n_scales=1, kind="kitti"):
    """Inputs of the loss path alone: images, intrinsics, depth pyramids, poses."""
    g = torch.Generator().manual_seed(seed)
    tgt = image(g, B, H, W)
    refs = [image(g, B, H, W) for _ in range(n_ref)]
    K = intrinsics(B, H, W, kind)
    tgt_depth = [depth_map(g, B, H >> s, W >> s) for s in range(n_scales)]
    ref_depths = [[depth_map(g, B, H >> s, W >> s) for s in range(n_scales)] for _ in range(n_ref)]
    poses = [pose(g, B) for _ in range(n_ref)]
    poses_inv = [pose(g, B) for _ in range(n_ref)]
    return dict(tgt_img=tgt, ref_imgs=refs, intrinsics=K, tgt_depth=tgt_depth, ref_depths=ref_depths,
                poses=poses, poses_inv=poses_inv)
