"""CPU oracle for the training-time image transforms.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy, what the reference's transform chain does to one sample (a list of H x W x 3 images that share one set
of random draws, plus the 3 x 3 intrinsics):

    RandomHorizontalFlip  custom_transforms.py:46-60   flip every image left-right, cx <- W - cx
    RandomScaleCrop       custom_transforms.py:63-89   zoom by (sx, sy) in [1, 1.15)^2 with PIL's Image.resize on the uint8
                                                       image, crop back to H x W at a random offset, scale / shift K
    ArrayToTensor         custom_transforms.py:32-43   HWC -> CHW, float32, / 255
    Normalize             custom_transforms.py:21-29   (x - 0.45) / 0.225 per channel, float32, in place

Third-party arithmetic: `Image.resize((w, h))` is Pillow's (not vendored by the reference; requirements.txt does not pin it;
installed here: Pillow 12.2.0).  For RGB images its default filter is BICUBIC, implemented in src/libImaging/Resample.c as two
separable passes (horizontal, then vertical) over 8-bit data with 22-bit fixed-point coefficients and an 8-bit intermediate
image.  `resize_bicubic_u8` restates that algorithm; tests/test_augment_cpu.py pins it bit for bit against the installed Pillow
and the whole chain against the unmodified reference classes (golden vectors: tests/golden/augment.npz, written by
tests/golden/make_golden_augment.py).
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are rounded to 22 fractional bits


def _bicubic(x):
    """Resample.c bicubic_filter, a = -0.5 (Keys), evaluated in double."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size): per output coordinate the first input
    coordinate, the tap count and the integer taps (<= 2 * ceil(support) + 1 of them)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img, out_w, out_h):
    """Image.fromarray(img).resize((out_w, out_h)) for an H x W x C uint8 array: horizontal pass into an 8-bit intermediate,
    then the vertical pass (Resample.c ImagingResampleInner; a pass whose size does not change is skipped there, which the
    exact 0 / 1 taps of the same-size case reproduce)."""
    img = np.asarray(img, np.uint8)
    H, W, C = img.shape
    bx, kx = resample_coeffs(W, out_w)
    by, ky = resample_coeffs(H, out_h)
    tmp = np.zeros((H, out_w, C), np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_w):
        x0, n = bx[xx]
        acc = np.full((H, C), 1 << (PRECISION_BITS - 1), np.int64)
        for i in range(n):
            acc += src[:, x0 + i, :] * int(kx[xx, i])
        tmp[:, xx, :] = _clip8(acc)
    out = np.zeros((out_h, out_w, C), np.uint8)
    src = tmp.astype(np.int64)
    for yy in range(out_h):
        y0, n = by[yy]
        acc = np.full((out_w, C), 1 << (PRECISION_BITS - 1), np.int64)
        for i in range(n):
            acc += src[y0 + i, :, :] * int(ky[yy, i])
        out[yy] = _clip8(acc)
    return out


def transform_sample(images, intrinsics, flip, x_scaling, y_scaling, offset_x, offset_y,
                     mean=(0.45, 0.45, 0.45), std=(0.225, 0.225, 0.225)):
    """The train-time chain on one sample with the random draws given: `flip` (random.random() < 0.5), the two zoom
    factors (np.random.uniform(1, 1.15, 2): x first) and the crop offsets (np.random.randint(scaled - in + 1): y first in
    the reference's draw order).  Returns (list of 3 x H x W float32 arrays, 3 x 3 intrinsics of the input dtype)."""
    K = np.copy(intrinsics)
    imgs = [np.asarray(im) for im in images]
    in_h, in_w, _ = imgs[0].shape
    if flip:
        imgs = [np.copy(np.fliplr(im)) for im in imgs]
        K[0, 2] = in_w - K[0, 2]
    scaled_h, scaled_w = int(in_h * y_scaling), int(in_w * x_scaling)
    K[0] *= x_scaling
    K[1] *= y_scaling
    out = []
    for im in imgs:
        big = resize_bicubic_u8(im.astype(np.uint8), scaled_w, scaled_h).astype(np.float32)
        crop = big[offset_y:offset_y + in_h, offset_x:offset_x + in_w]
        t = np.transpose(crop, (2, 0, 1)).astype(np.float32) / np.float32(255)
        for c in range(t.shape[0]):
            t[c] = (t[c] - np.float32(mean[c])) / np.float32(std[c])
        out.append(t)
    K[0, 2] -= offset_x
    K[1, 2] -= offset_y
    return out, K


def plain_sample(images, mean=(0.45, 0.45, 0.45), std=(0.225, 0.225, 0.225)):
    """The validation chain (ArrayToTensor + Normalize only)."""
    out = []
    for im in images:
        t = np.transpose(np.asarray(im, np.float32), (2, 0, 1)) / np.float32(255)
        for c in range(t.shape[0]):
            t[c] = (t[c] - np.float32(mean[c])) / np.float32(std[c])
        out.append(t)
    return out
