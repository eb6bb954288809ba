"""One Generator forward+backward at batch 64 (after one warm-up pass) for ncu captures:
  ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc' -s 65 -c 65 \
      -o gpurun_out/prof python tools/profile_g.py [parity|fast] [B]
(65 tensor-core launches per pass: 20 fwd convs, 25 dgrad convs, 20 wgrads)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcgvc_loader  # noqa: E402

pkg = mcgvc_loader.load()
eng = pkg.engine
mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "fast": eng.PRECISION_FAST, "mixed": eng.PRECISION_MIXED}[mode])
torch.manual_seed(0)
G = pkg.Generator().to("cuda")
x = torch.randn(B, 80, 64, device="cuda")
m = torch.ones_like(x)
for it in range(2):
    G.zero_grad(set_to_none=True)
    y = G(x, m)
    y.abs().mean().backward()
    torch.cuda.synchronize()
print("done", float(y.abs().mean()))
