"""One Generator + Discriminator adversarial forward+backward at batch 64 (after one warm-up pass):
    ncu --metrics gpu__time_duration.sum --clock-control none -s <n> -c <m> --csv python tools/profile_gd.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcgvc_loader  # noqa: E402

pkg = mcgvc_loader.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
G, D = pkg.Generator().to("cuda"), pkg.Discriminator().to("cuda")
x = torch.randn(B, 80, 64, device="cuda", requires_grad=True)
m = torch.ones(B, 80, 64, device="cuda")
for it in range(2):
    G.zero_grad(set_to_none=True)
    D.zero_grad(set_to_none=True)
    loss = torch.mean((1 - D(G(x, m))) ** 2)
    loss.backward()
    torch.cuda.synchronize()
    if it == 0:
        import ctypes
        print("launches in one pass:", pkg.engine.launch_count())
print("done", float(loss))
