"""TEST INFRASTRUCTURE: synthetic stand-in for the preprocessed VCC2018 data the reference's drivers
load (train.py:51-64, test.py:41-55; format written by data_preprocessing/preprocess_vcc2018.py:
`<dir>/<spk>/<spk>_normalized.pickle` = list of (80, T_i) float arrays, `<spk>_norm_stat.npz` with
`mean`, `std` of shape (80, 1)).  Utterance lengths are deliberately not multiples of 4."""
import os
import pickle
import sys

import numpy as np


def make(root, speakers=("SPKA", "SPKB"), n_utts=4, seed=7, lengths=(101, 97, 135, 118, 90, 143)):
    rng = np.random.RandomState(seed)
    for s_i, spk in enumerate(speakers):
        d = os.path.join(root, spk)
        os.makedirs(d, exist_ok=True)
        utts = []
        for u in range(n_utts):
            T = lengths[(u + 2 * s_i) % len(lengths)]
            # smooth-ish standardised mel: AR(1) along time, per-bin offsets
            e = rng.randn(80, T).astype(np.float32)
            x = np.zeros_like(e)
            x[:, 0] = e[:, 0]
            for t in range(1, T):
                x[:, t] = 0.8 * x[:, t - 1] + 0.6 * e[:, t]
            utts.append(x.astype(np.float32))
        with open(os.path.join(d, "%s_normalized.pickle" % spk), "wb") as f:
            pickle.dump(utts, f)
        np.savez(os.path.join(d, "%s_norm_stat.npz" % spk),
                 mean=(0.1 * rng.randn(80, 1)).astype(np.float32), std=(1.0 + 0.1 * np.abs(rng.randn(80, 1))).astype(np.float32))
    return root


if __name__ == "__main__":
    make(sys.argv[1])
    print("synthetic data written under", sys.argv[1])
