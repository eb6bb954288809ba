"""Device-side data feed: host sampler + crop/mask kernel against the oracle
restatement of dataset/vc_dataset.py:44-56 and, when /root/reference is present, against the
reference's own VCDataset driven by the same random draws."""
import os
import sys

import numpy as np
import pytest
import torch

import maskcyclegan_oracle as O


def _datasets(seed=0, n=7):
    rng = np.random.RandomState(seed)
    A = [rng.randn(80, int(t)).astype(np.float32) for t in rng.randint(64, 200, size=n)]
    B = [rng.randn(80, int(t)).astype(np.float32) for t in rng.randint(64, 200, size=n + 2)]
    return A, B


def test_selection_follows_the_reference_distributions(pkg):
    frames = np.array([64, 100, 150], dtype=np.int32)
    sel = pkg.draw_selection(np.random.RandomState(1), frames, 4000, n_frames=64, max_mask_len=25)
    u, start, mstart, msize = sel
    assert set(np.unique(u)) == {0, 1, 2}
    assert (start >= 0).all() and (start + 64 <= frames[u]).all() and (start[u == 0] == 0).all()
    assert msize.min() == 0 and msize.max() == 24                      # U{0 .. max_mask_len - 1}
    assert (mstart >= 0).all() and (mstart + msize <= 63).all()        # U{0 .. n_frames - size - 1}
    assert abs(np.mean(u == 1) - 1 / 3) < 0.04 and abs(msize.mean() - 12.0) < 0.6
    with pytest.raises(ValueError):
        pkg.draw_selection(np.random.RandomState(0), np.array([63]), 1)


def test_oracle_crop_matches_the_reference_dataset_when_present():
    """Replays the reference's own __getitem__ (vc_dataset.py:19-77) with a recording RNG and checks the
    oracle restatement reproduces its outputs from the recorded draws."""
    if not os.path.isdir("/root/reference/dataset"):
        pytest.skip("reference checkout not present (GPU box)")
    sys.path.insert(0, "/root/reference")
    try:
        from dataset.vc_dataset import VCDataset
    finally:
        sys.path.remove("/root/reference")
    A, B = _datasets(3, n=5)
    ds = VCDataset(A, B, n_frames=64, max_mask_len=25)
    for index in (0, 3):
        np.random.seed(100 + index)
        xa, ma, xb, mb = ds[index]
        # the same draws, replayed: two permutations, then (start, size, mask start) for A and B per pair
        np.random.seed(100 + index)
        ia, ib = np.arange(len(A)), np.arange(len(B))
        np.random.shuffle(ia)
        np.random.shuffle(ib)
        n = min(len(A), len(B))
        selA, selB = [], []
        for ua, ub in zip(ia[:n], ib[:n]):
            sa = np.random.randint(A[ua].shape[1] - 64 + 1)
            za = np.random.randint(0, 25)
            selA.append((ua, sa, np.random.randint(0, 64 - za), za))
            sb = np.random.randint(B[ub].shape[1] - 64 + 1)
            zb = np.random.randint(0, 25)
            selB.append((ub, sb, np.random.randint(0, 64 - zb), zb))
        oa, oma = O.crop_and_mask(A, np.array(selA, dtype=np.int32).T[:, index:index + 1])
        ob, omb = O.crop_and_mask(B, np.array(selB, dtype=np.int32).T[:, index:index + 1])
        assert np.array_equal(oa[0], xa) and np.array_equal(oma[0], ma)
        assert np.array_equal(ob[0], xb) and np.array_equal(omb[0], mb)


@pytest.mark.gpu
def test_crop_mask_kernel_is_bit_exact_against_the_oracle(pkg):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    A, B = _datasets(5)
    feed = pkg.DeviceVCDataFeed(A, B, batch_size=16, n_frames=64, max_mask_len=25, device="cuda", seed=7)
    for pool, data in ((feed.A, A), (feed.B, B)):
        sel = pkg.draw_selection(np.random.RandomState(11), pool.frames, 16)
        sel[:, 0] = (0, 0, 0, 0)                                   # empty mask at the very start
        sel[:, 1] = (1, pool.frames[1] - 64, 63 - 24, 24)          # last possible crop, longest mask at the end
        x, m = feed.crop(pool, sel)
        ox, om = O.crop_and_mask(data, sel)
        assert torch.equal(x.cpu(), torch.from_numpy(ox)) and torch.equal(m.cpu(), torch.from_numpy(om))
    xa, ma, xb, mb = feed.next_batch()
    assert xa.shape == ma.shape == xb.shape == mb.shape == (16, 80, 64) and xa.is_cuda
    assert set(torch.unique(ma).tolist()) <= {0.0, 1.0} and len(feed) == 1
    # other frame counts, and an out-of-range selection never reads outside the pool
    feed2 = pkg.DeviceVCDataFeed(A, B, batch_size=3, n_frames=40, max_mask_len=10, device="cuda", seed=1)
    sel = pkg.draw_selection(np.random.RandomState(2), feed2.A.frames, 3, n_frames=40, max_mask_len=10)
    x, m = feed2.crop(feed2.A, sel)
    ox, om = O.crop_and_mask(A, sel, n_frames=40)
    assert torch.equal(x.cpu(), torch.from_numpy(ox)) and torch.equal(m.cpu(), torch.from_numpy(om))
    bad = sel.copy()
    bad[0, 0], bad[1, 1] = 99, 10 ** 6
    x, m = feed2.crop(feed2.A, bad)
    assert torch.count_nonzero(x[0]) == 0 and torch.count_nonzero(x[1]) == 0 and torch.equal(x[2].cpu(), torch.from_numpy(ox[2]))
