"""Device-side training transforms (csrc/augment.cu through scsfm.augment.GpuAugment) against the oracle (oracle/augment.py,
itself pinned bit for bit against the unmodified reference chain and Pillow in tests/test_augment_cpu.py) and the golden
vectors of the reference chain.  Byte arithmetic + three IEEE float32 operations: the comparison is exact equality."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import augment as A

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment.npz")


def _aug():
    from scsfm import augment
    return augment


@pytest.mark.parametrize("seed", (0, 1, 2, 3, 5, 8))
def test_golden_vectors_of_the_reference_chain(seed):
    G = _aug()
    gold = np.load(GOLD)
    p = "s%d_" % seed
    flip, xs, ys, ox, oy = gold[p + "draws"]
    imgs = gold[p + "images"]                                  # [n,H,W,3] uint8
    n, H, W, _ = imgs.shape
    d = G.Draw.given(H, W, bool(flip), xs, ys, int(ox), int(oy))
    out, K = G.GpuAugment()(torch.from_numpy(imgs)[:, None], gold[p + "K"][None], draws=[d])
    got = torch.stack(out)[:, 0].cpu().numpy()
    assert np.array_equal(got, gold[p + "out"])
    assert np.array_equal(K[0].cpu().numpy(), gold[p + "K_out"])
    plain, K0 = G.GpuAugment(train=False)(torch.from_numpy(imgs)[:, None].float(), gold[p + "K"][None])     # float frames, like load_as_float
    assert np.array_equal(torch.stack(plain)[:, 0].cpu().numpy(), gold[p + "plain"])
    assert np.array_equal(K0[0].cpu().numpy(), gold[p + "K"])


@pytest.mark.parametrize("shape", [(4, 3, 256, 832), (8, 2, 256, 320), (2, 3, 37, 53)])
def test_full_size_batch_equals_oracle_and_follows_the_reference_draw_order(shape):
    G = _aug()
    B, n, H, W = shape
    g = np.random.default_rng(7)
    low = g.integers(0, 256, (n, B, H // 8 + 2, W // 8 + 2, 3)).astype(np.float32)
    imgs = np.kron(low, np.ones((1, 1, 8, 8, 1), np.float32))[:, :, :H, :W]
    imgs = np.clip(imgs + g.normal(0, 25, imgs.shape), 0, 255).astype(np.uint8)
    K = np.tile(np.array([[0.58 * W, 0, 0.49 * W], [0, 1.92 * H, 0.47 * H], [0, 0, 1]], np.float32), (B, 1, 1))
    K[:, 0, 2] += np.arange(B, dtype=np.float32)
    random.seed(3)
    np.random.seed(3)
    out, K_out = G.GpuAugment()(torch.from_numpy(imgs).pin_memory(), K)
    torch.cuda.synchronize()
    # the oracle with the draws taken in the reference's order from the same RNG state
    random.seed(3)
    np.random.seed(3)
    flips = 0
    for b in range(B):
        flip = random.random() < 0.5
        xs, ys = np.random.uniform(1, 1.15, 2)
        oy = np.random.randint(int(H * ys) - H + 1)
        ox = np.random.randint(int(W * xs) - W + 1)
        flips += flip
        ref, Kb = A.transform_sample([imgs[i, b].astype(np.float32) for i in range(n)], K[b], flip, xs, ys, ox, oy)
        for i in range(n):
            assert np.array_equal(out[i][b].cpu().numpy(), ref[i]), (b, i)
        assert np.array_equal(K_out[b].cpu().numpy(), Kb)
    assert out[0].shape == (B, 3, H, W) and out[0].is_contiguous() and K_out.dtype == torch.float32
    if B >= 4:
        assert 0 < flips < B            # both branches of the flip were exercised


def test_error_behaviour():
    G = _aug()
    aug = G.GpuAugment()
    with pytest.raises(ValueError):
        aug(torch.zeros(2, 1, 8, 8, 4, dtype=torch.uint8), np.zeros((1, 3, 3), np.float32))
    with pytest.raises(ValueError):
        aug(torch.zeros(2, 1, 8, 8, 3, dtype=torch.uint8), np.zeros((2, 3, 3), np.float32))
    with pytest.raises(ValueError):
        aug(torch.zeros(2, 1, 8, 8, 3, dtype=torch.uint8), np.zeros((1, 3, 3), np.float32), draws=[G.Draw(False, 1.0, 1.0, 8, 8, 1, 0)])


def test_loader_path_of_a_real_dataset():
    """train.py's RawFrames + DataLoader + GpuAugmentLoader on a stand-in for the reference's SequenceFolder(transform=None)
    (float32 H x W x 3 frames with integer values, float32 intrinsics and their inverse): the loop's tuple comes out on the
    GPU, equal to the oracle's validation chain (train=False) and, with train=True, to the oracle under the same draws."""
    import train as T
    H, W, n = 24, 40, 5
    g = np.random.default_rng(1)
    frames = g.integers(0, 256, (n, 3, H, W, 3)).astype(np.float32)
    K = np.array([[0.58 * W, 0, 0.49 * W], [0, 1.92 * H, 0.47 * H], [0, 0, 1]], np.float32)

    class Fake(torch.utils.data.Dataset):
        def __len__(self):
            return n

        def __getitem__(self, i):
            return frames[i, 0], [frames[i, 1], frames[i, 2]], np.copy(K), np.linalg.inv(K)

    loader = torch.utils.data.DataLoader(T.RawFrames(Fake()), batch_size=2, shuffle=False, drop_last=True)
    dev = torch.device("cuda")
    batches = list(T.GpuAugmentLoader(loader, dev, train=False))
    assert len(batches) == 2
    for bi, (tgt, refs, Kb, Kinv) in enumerate(batches):
        assert tgt.shape == (2, 3, H, W) and len(refs) == 2 and tgt.is_cuda and Kb.shape == (2, 3, 3)
        for j in range(2):
            want = A.plain_sample([frames[2 * bi + j, i] for i in range(3)])
            got = [tgt[j]] + [r[j] for r in refs]
            for i in range(3):
                assert np.array_equal(got[i].cpu().numpy(), want[i])
        torch.testing.assert_close(torch.bmm(Kb, Kinv), torch.eye(3, device=dev).expand(2, 3, 3), atol=1e-4, rtol=0)
    random.seed(11)
    np.random.seed(11)
    tgt, refs, Kb, _ = next(iter(T.GpuAugmentLoader(loader, dev, train=True)))
    random.seed(11)
    np.random.seed(11)
    G = _aug()
    for j in range(2):
        d = G.Draw.random(H, W)
        want, Kw = A.transform_sample([frames[j, i] for i in range(3)], K, d.flip, d.x_scaling, d.y_scaling, d.offset_x, d.offset_y)
        got = [tgt[j]] + [r[j] for r in refs]
        for i in range(3):
            assert np.array_equal(got[i].cpu().numpy(), want[i])
        assert np.array_equal(Kb[j].cpu().numpy(), Kw)
