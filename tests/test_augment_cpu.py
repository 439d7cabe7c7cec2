"""The oracle of the training-time image transforms (oracle/augment.py) against the unmodified reference chain (golden
vectors written by tests/golden/make_golden_augment.py) and its restatement of Pillow's bicubic resize against the
installed Pillow.  Bit-exact: the pipeline is byte arithmetic followed by three IEEE float32 operations per value."""
import os

import numpy as np
import pytest

from oracle import augment as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment.npz")
SEEDS = (0, 1, 2, 3, 5, 8)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_chain_equals_reference_chain(gold, seed):
    p = "s%d_" % seed
    flip, xs, ys, ox, oy = gold[p + "draws"]
    imgs = [im.astype(np.float32) for im in gold[p + "images"]]
    out, K = A.transform_sample(imgs, gold[p + "K"], bool(flip), xs, ys, int(ox), int(oy))
    assert np.array_equal(np.stack(out), gold[p + "out"])
    assert K.dtype == gold[p + "K_out"].dtype and np.array_equal(K, gold[p + "K_out"])
    assert np.array_equal(np.stack(A.plain_sample(imgs)), gold[p + "plain"])


def test_bicubic_restatement_equals_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for H, W in [(24, 40), (37, 53), (64, 208)]:
        for t in range(4):
            img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            sx, sy = rng.uniform(1, 1.15, 2)
            ow, oh = int(W * sx), int(H * sy)
            if t == 0:
                ow, oh = W, H               # identity: both passes skipped by Pillow, exact 0 / 1 taps here
            if t == 1:
                ow = W                      # vertical pass only
            ref = np.array(Image.fromarray(img).resize((ow, oh)))
            assert np.array_equal(A.resize_bicubic_u8(img, ow, oh), ref), (H, W, ow, oh)


def test_coefficients_are_normalised_and_bounded():
    for n_in, n_out in [(256, 256), (256, 294), (832, 956), (10, 11)]:
        b, k = A.resample_coeffs(n_in, n_out)
        assert k.shape[1] == 5 and (b[:, 1] <= 5).all() and (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= n_in).all()
        assert np.abs(k.sum(1) - (1 << A.PRECISION_BITS)).max() <= 3          # rounding of <= 5 taps


def _host():
    import importlib
    return importlib.import_module("scsfm.augment")


@pytest.mark.parametrize("seed", SEEDS)
def test_host_side_draw_order_and_intrinsics_follow_the_reference(gold, seed):
    """scsfm.augment.Draw.random consumes the global RNGs in the reference's order and update_intrinsics reproduces the
    reference's K (host logic of the device pipeline; the kernels are checked in tests/test_augment_gpu.py)."""
    import random
    G = _host()
    p = "s%d_" % seed
    _, H, W, _ = gold[p + "images"].shape
    random.seed(seed)
    np.random.seed(seed)
    d = G.Draw.random(H, W)
    flip, xs, ys, ox, oy = gold[p + "draws"]
    assert (d.flip, d.x_scaling, d.y_scaling, d.offset_x, d.offset_y) == (bool(flip), xs, ys, int(ox), int(oy))
    assert np.array_equal(G.update_intrinsics(gold[p + "K"], d, W), gold[p + "K_out"])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    G = _host()
    with pytest.raises(RuntimeError):
        G.GpuAugment()(np.zeros((1, 1, 8, 8, 3), np.uint8), np.zeros((1, 3, 3), np.float32))
