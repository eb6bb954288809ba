"""One warm Generator + Discriminator adversarial forward + backward at batch B, then one fused-Adam step
and the re-pack it triggers, bracketed by cudaProfilerStart/Stop -- the capture target for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o <rep> \
        python tools/profile_pass.py [c8|c8h|parity] [B]
(every kernel of the hot path appears once per use: convs, wgrads, fused trunk, layer kernels, stems,
heads, unpack, Adam, pack)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcgvc_loader  # noqa: E402

pkg = mcgvc_loader.load()
eng = pkg.engine
mode = sys.argv[1] if len(sys.argv) > 1 else "c8"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "c8w": eng.PRECISION_C8W, "c8h": eng.PRECISION_C8H,
                   "mixed": eng.PRECISION_MIXED, "fast": eng.PRECISION_FAST}[mode])
torch.manual_seed(0)
G, D = pkg.Generator().to("cuda"), pkg.Discriminator().to("cuda")
opt = pkg.FusedAdam([G, D], lr=2e-4, betas=(0.5, 0.999))
x = torch.randn(B, 80, 64, device="cuda", requires_grad=True)
m = torch.ones(B, 80, 64, device="cuda")


def one_pass():
    G.zero_grad(set_to_none=True)
    D.zero_grad(set_to_none=True)
    loss = torch.mean((1 - D(G(x, m))) ** 2)
    loss.backward()
    opt.step()
    with torch.no_grad():
        G(x, m)          # re-packs the updated weights
    return loss


for _ in range(2):
    one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(loss))
