"""Device-side training data feed.

Host-side mirror of the reference's `VCDataset` (dataset/vc_dataset.py:12-82) + the `.to(device)`
copies of mask_cyclegan_vc/train.py:187-190, for training only (`valid=False`): the two mel
datasets are uploaded ONCE into device-resident pools; every batch then costs one 16-byte-per-sample
selection upload and one `mcgvc_crop_mask` launch per side instead of re-shuffling and re-cropping
the whole dataset on the host for every item and copying 2 x 80 x n_frames floats per sample.

Sampling semantics of the reference (vc_dataset.py:32-56), per returned sample and per side:
utterance ~ uniform over the dataset (each `__getitem__` re-shuffles, so items are independent),
crop start ~ U{0 .. T_u - n_frames}, mask size ~ U{0 .. max_mask_len - 1}, mask start ~
U{0 .. n_frames - mask_size - 1}.  The draws come from a `numpy.random.RandomState` owned by the
feed (the reference uses numpy's global state; its exact stream -- O(dataset) draws per item --
is not reproduced, the distributions are).
"""
import ctypes

import numpy as np
import torch

from . import engine


def draw_selection(rng, frames, batch, n_frames=64, max_mask_len=25):
    """int32 [4, batch]: utterance, crop start, mask start, mask size (vc_dataset.py:44-52)."""
    frames = np.asarray(frames)
    if frames.min() < n_frames:
        raise ValueError("every utterance needs at least n_frames=%d frames (vc_dataset.py:47)" % n_frames)
    if not n_frames > max_mask_len - 1:
        raise ValueError("n_frames must exceed the largest mask size (vc_dataset.py:51)")
    sel = np.empty((4, batch), dtype=np.int32)
    for b in range(batch):
        u = rng.randint(len(frames))
        sel[0, b] = u
        sel[1, b] = rng.randint(frames[u] - n_frames + 1)
        size = rng.randint(0, max_mask_len)
        sel[3, b] = size
        sel[2, b] = rng.randint(0, n_frames - size)
    return sel


class _Pool:
    def __init__(self, dataset, device):
        arrays = [np.ascontiguousarray(np.asarray(a), dtype=np.float32) for a in dataset]
        for a in arrays:
            if a.ndim != 2 or a.shape[0] != 80:
                raise ValueError("utterances must be (80, T) mel arrays, got %s" % (a.shape,))
        self.frames = np.array([a.shape[1] for a in arrays], dtype=np.int32)
        off = np.zeros(len(arrays), dtype=np.int64)
        off[1:] = np.cumsum([a.size for a in arrays[:-1]])
        self.pool = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrays])).to(device)
        self.off = torch.from_numpy(off).to(device)
        self.frames_dev = torch.from_numpy(self.frames).to(device)


class DeviceVCDataFeed:
    """`for real_A, mask_A, real_B, mask_B in feed:` yields device tensors of shape (B, 80, n_frames)."""

    def __init__(self, datasetA, datasetB, batch_size, n_frames=64, max_mask_len=25, device="cuda", seed=0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise engine.EngineError("DeviceVCDataFeed lives on a CUDA device (no CPU path)")
        self.n_frames, self.max_mask_len, self.batch_size = int(n_frames), int(max_mask_len), int(batch_size)
        self.A, self.B = _Pool(datasetA, self.device), _Pool(datasetB, self.device)
        self.rng = np.random.RandomState(seed)
        self._len = min(len(self.A.frames), len(self.B.frames))      # vc_dataset.py:78-82

    def __len__(self):
        return (self._len + self.batch_size - 1) // self.batch_size  # batches per "epoch" like the DataLoader

    def crop(self, pool, sel):
        """One side: selection [4, B] (host int32) -> (x, mask) on the device."""
        B = sel.shape[1]
        sel_dev = torch.from_numpy(np.ascontiguousarray(sel, dtype=np.int32)).to(self.device, non_blocking=True)
        x = torch.empty(B, 80, self.n_frames, dtype=torch.float32, device=self.device)
        mask = torch.empty_like(x)
        lib = engine.lib()
        lib.mcgvc_set_device(self.device.index or 0)
        rc = lib.mcgvc_crop_mask(ctypes.c_void_p(pool.pool.data_ptr()), ctypes.c_void_p(pool.off.data_ptr()),
                                 ctypes.c_void_p(pool.frames_dev.data_ptr()), len(pool.frames),
                                 ctypes.c_void_p(sel_dev.data_ptr()), B, self.n_frames,
                                 ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(mask.data_ptr()), ctypes.c_void_p(0),
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise engine.EngineError("crop_mask failed: %s" % lib.mcgvc_last_error().decode())
        return x, mask

    def next_batch(self):
        selA = draw_selection(self.rng, self.A.frames, self.batch_size, self.n_frames, self.max_mask_len)
        selB = draw_selection(self.rng, self.B.frames, self.batch_size, self.n_frames, self.max_mask_len)
        xa, ma = self.crop(self.A, selA)
        xb, mb = self.crop(self.B, selB)
        return xa, ma, xb, mb

    def __iter__(self):
        for _ in range(len(self)):
            yield self.next_batch()
