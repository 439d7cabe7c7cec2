"""Deterministic weights / images shared by make_golden.py and the tests.

numpy's legacy MT19937 RandomState is bit-stable across numpy versions and platforms,
so keying it by crc32(parameter name) reproduces the exact tensors anywhere without
relying on torch's RNG stream or on module construction order.
"""
import zlib

import numpy as np
import torch


def _rs(name):
    return np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)


def det_weights(state_dict):
    """Fill every entry of a DispResNet/PoseResNet state_dict deterministically."""
    out = {}
    for k, v in state_dict.items():
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_mean"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            out[k] = torch.ones_like(v)
        else:
            n = _rs(k).standard_normal(shape).astype(np.float32)
            if v.dim() == 4:
                fan_in = shape[1] * shape[2] * shape[3]
                n *= np.float32(np.sqrt(2.0 / fan_in))
            elif v.dim() == 2:
                n *= np.float32(0.01)
            elif k.endswith(".weight"):
                n = np.float32(1.0) + np.float32(0.1) * n
            else:
                n *= np.float32(0.1)
            out[k] = torch.from_numpy(n)
    return out


def det_image(name, B, H, W):
    """Smooth pseudo-image in the normalised range of the dataset transform."""
    u = _rs(name).uniform(0, 1, (B, 3, H + 4, W + 4)).astype(np.float32)
    t = torch.from_numpy(u)
    t = torch.nn.functional.avg_pool2d(t, 5, 1)
    return ((t - 0.45) / 0.225).contiguous()
