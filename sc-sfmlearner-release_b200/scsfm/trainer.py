"""One optimisation step of the reference training loop (reference train.py:259-282, 426-444) on one
B200, and its data-parallel form: one process per GPU, each rank runs the complete step on its shard of
the batch, the two gradient arenas are all-reduced (NCCL over NVLink) as soon as the backward of the
owning network has finished, then every rank applies the identical fused Adam update.
"""
import torch
import torch.distributed as dist

import loss_functions as LF

from .exchange import GradExchange
from .nets import ArenaAdam


def compute_depth(disp_net, tgt_img, ref_imgs):
    """train.py:426-434 -- depth = 1/disparity for the target and every reference, one network call each
    (BatchNorm statistics stay per call)."""
    tgt_depth = [1 / d for d in disp_net(tgt_img)]
    ref_depths = [[1 / d for d in disp_net(r)] for r in ref_imgs]
    return tgt_depth, ref_depths


def compute_pose_with_inv(pose_net, tgt_img, ref_imgs):
    """train.py:437-444."""
    poses = [pose_net(tgt_img, r) for r in ref_imgs]
    poses_inv = [pose_net(r, tgt_img) for r in ref_imgs]
    return poses, poses_inv


class Trainer:
    """Holds the two networks, the fused Adam and (optionally) the data-parallel gradient exchange."""

    def __init__(self, disp_net, pose_net, lr=1e-4, betas=(0.9, 0.999), weight_decay=0.0, num_scales=1, with_ssim=1,
                 with_mask=1, with_auto_mask=0, padding_mode="zeros", w1=1.0, w2=0.1, w3=0.5, distributed=None):
        self.disp_net, self.pose_net = disp_net, pose_net
        self.optimizer = ArenaAdam([disp_net, pose_net], lr=lr, betas=betas, weight_decay=weight_decay)
        self.cfg = dict(num_scales=num_scales, with_ssim=with_ssim, with_mask=with_mask, with_auto_mask=with_auto_mask,
                        padding_mode=padding_mode)
        self.w = (w1, w2, w3)
        self.distributed = dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size() if self.distributed else 1
        self.exchange = None
        if self.distributed:
            self.exchange = GradExchange(self.world, torch.cuda.Stream())
            for net in (disp_net, pose_net):
                net.ensure_arena()
                # fires when the last pending backward of the network has been enqueued: its gradient arena is
                # averaged on the side stream while the rest of the backward pass keeps running
                net.grads_ready_callback = lambda n: self.exchange.allreduce_async(n.flat_grads())
            self.exchange.broadcast_params([disp_net.flat_params(), pose_net.flat_params()])   # identical replicas

    def losses(self, tgt_img, ref_imgs, intrinsics):
        tgt_depth, ref_depths = compute_depth(self.disp_net, tgt_img, ref_imgs)
        poses, poses_inv = compute_pose_with_inv(self.pose_net, tgt_img, ref_imgs)
        c = self.cfg
        photo, geo = LF.compute_photo_and_geometry_loss(tgt_img, ref_imgs, intrinsics, tgt_depth, ref_depths, poses,
                                                        poses_inv, c["num_scales"], c["with_ssim"], c["with_mask"],
                                                        c["with_auto_mask"], c["padding_mode"])
        smooth = LF.compute_smooth_loss(tgt_depth, tgt_img, ref_depths, ref_imgs)
        w1, w2, w3 = self.w
        return w1 * photo + w2 * smooth + w3 * geo, photo, smooth, geo

    def step(self, tgt_img, ref_imgs, intrinsics):
        """train.py:259-282 without a single host synchronisation; returns device scalars
        (loss, photo, smooth, geometry)."""
        loss, photo, smooth, geo = self.losses(tgt_img, ref_imgs, intrinsics)
        self.optimizer.zero_grad()
        loss.backward()
        if self.exchange is not None:
            self.exchange.wait()
        self.optimizer.step()
        return loss.detach(), photo.detach(), smooth.detach(), geo.detach()
