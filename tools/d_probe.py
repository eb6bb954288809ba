"""Per-kernel timeline of one warm Discriminator forward + backward (input requiring grad) at batch 64."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import mcgvc_loader  # noqa: E402
from g_timeline import capture, table  # noqa: E402

pkg = mcgvc_loader.load()
eng = pkg.engine
mode = sys.argv[1] if len(sys.argv) > 1 else "c8w"
eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "c8w": eng.PRECISION_C8W, "c8h": eng.PRECISION_C8H}[mode])
torch.manual_seed(0)
D = pkg.Discriminator().cuda()
x = torch.randn(64, 80, 64, device="cuda", requires_grad=True)


def fb():
    D.zero_grad(set_to_none=True)
    x.grad = None
    ((1 - D(x)) ** 2).mean().backward()


print(table("Discriminator forward + backward, batch 64, %s" % mode, capture(fb)))
