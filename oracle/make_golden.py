"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from
/root/reference/mask_cyclegan_vc/model.py) on CPU fp32.  Run in the build container only:

    python oracle/make_golden.py

The fixtures pin oracle/maskcyclegan_oracle.py (tests/test_oracle.py) and, through it, the CUDA
engine (tests/test_gpu_*.py).  Weights are never stored: `torch.manual_seed(seed)` followed by the
reference's constructors is reproducible, and the fixtures carry per-tensor checksums so that the
replayed initialisation is itself verified on the GPU box (where /root/reference does not exist).
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

from mask_cyclegan_vc.model import Discriminator, Generator  # noqa: E402  (the reference)
import maskcyclegan_oracle as O  # noqa: E402

SAMPLES_PER_TENSOR = 8


def sample_idx(numel, k=SAMPLES_PER_TENSOR):
    # fixed, deterministic positions spread over the tensor
    return np.unique(np.linspace(0, numel - 1, k).astype(np.int64))


def tensor_digest(t):
    f = t.detach().double().flatten()
    idx = sample_idx(f.numel())
    return np.concatenate([[f.sum().item(), f.norm().item()], f[idx].numpy()])


def params_digest(module):
    # de-duplicated parameter order == optimizer order (train.py:113-122)
    return np.stack([np.pad(tensor_digest(p), (0, 2 + SAMPLES_PER_TENSOR - len(tensor_digest(p))))
                     for p in module.parameters()])


def grads_digest(module):
    rows = []
    for p in module.parameters():
        if p.grad is None:
            rows.append(np.full(2 + SAMPLES_PER_TENSOR, np.nan))
        else:
            d = tensor_digest(p.grad)
            rows.append(np.pad(d, (0, 2 + SAMPLES_PER_TENSOR - len(d))))
    return np.stack(rows)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    # ---- forward fixtures: Generator / Discriminator at several (B, T), seed 0 weights
    torch.manual_seed(0)
    G = Generator()
    D = Discriminator()
    np.savez_compressed(os.path.join(OUT, "weights_seed0.npz"),
                        g_digest=params_digest(G), d_digest=params_digest(D),
                        g_names=np.array([n for n, _ in G.named_parameters()]),
                        d_names=np.array([n for n, _ in D.named_parameters()]))
    for (B, T) in ((1, 64), (2, 64), (1, 65), (1, 100), (3, 32)):
        real_A, mask_A, _, _ = O.synthetic_batch(B, T, seed=1234 + T + B, max_mask_len=min(25, T // 2))
        with torch.no_grad():
            y = G(real_A, mask_A)
            y1 = G(real_A, torch.ones_like(real_A))
            d = D(real_A)
            dy = D(y)
        np.savez_compressed(os.path.join(OUT, "fwd_B%d_T%d.npz" % (B, T)), x=real_A.numpy(),
                            mask=mask_A.numpy(), g_out=y.numpy(), g_out_ones=y1.numpy(),
                            d_out=d.numpy(), d_of_g=dy.numpy())
        print("fwd", B, T, y.shape, d.shape)

    # ---- config 2: single adversarial fwd+bwd (BASELINE.json configs[1]), B=1 and B=2
    for B in (1, 2):
        G.zero_grad(set_to_none=True)
        D.zero_grad(set_to_none=True)
        real_A, mask_A, _, _ = O.synthetic_batch(B, 64, seed=77 + B)
        x = real_A.clone().requires_grad_(True)
        fake = G(x, mask_A)
        d = D(fake)
        loss = torch.mean((1 - d) ** 2)
        loss.backward()
        np.savez_compressed(os.path.join(OUT, "adv_B%d.npz" % B), x=real_A.numpy(), mask=mask_A.numpy(),
                            loss=np.float64(loss.item()), g_grads=grads_digest(G),
                            d_grads=grads_digest(D), x_grad=x.grad.numpy(), fake=fake.detach().numpy())
        print("adv", B, loss.item())

    # ---- two full train steps (train.py:186-299 semantics), B=2, T=64, seed 0 construction order
    torch.manual_seed(0)
    G_A2B, G_B2A = Generator(), Generator()
    D_A, D_B, D_A2, D_B2 = Discriminator(), Discriminator(), Discriminator(), Discriminator()
    g_opt = torch.optim.Adam(list(G_A2B.parameters()) + list(G_B2A.parameters()), lr=2e-4, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(list(D_A.parameters()) + list(D_B.parameters()) + list(D_A2.parameters()) +
                             list(D_B2.parameters()), lr=1e-4, betas=(0.5, 0.999))
    losses = []
    grad_digests = {}
    for step in range(2):
        batch = O.synthetic_batch(2, 64, seed=1234 + step)
        gl, dl = O.train_step(G_A2B, G_B2A, D_A, D_B, D_A2, D_B2, g_opt, d_opt, batch)
        losses.append((gl, dl))
        print("train step", step, gl, dl)
        if step == 0:
            # grads present after d_loss.backward(): D grads (used) and G grads (discarded later)
            grad_digests = {"d_A": grads_digest(D_A), "d_B2": grads_digest(D_B2), "g_A2B": grads_digest(G_A2B)}
    np.savez_compressed(os.path.join(OUT, "train_B2.npz"), losses=np.array(losses),
                        g_A2B_after=params_digest(G_A2B), g_B2A_after=params_digest(G_B2A),
                        d_A_after=params_digest(D_A), d_B2_after=params_digest(D_B2),
                        step0_d_A_grads=grad_digests["d_A"], step0_d_B2_grads=grad_digests["d_B2"],
                        step0_g_A2B_grads=grad_digests["g_A2B"])
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
