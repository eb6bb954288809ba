"""CPU estimate of what the C8W precision mode (fp16 single-pass weight-gradient GEMMs on the C8
layers) adds to the packed gradients, before any GPU time is spent on it.

Runs the reference's own modules (oracle/_ref, staged by oracle/build_ref.sh) on the CPU through a
G -> D adversarial pass, captures every C8 layer's input x and output gradient dz with hooks, and
recomputes that layer's weight gradient in fp64 from the operands as the engine's single-pass
kernels see them: x -> fp16(x), dz -> fp16(dz * S) / S with the engine's per-layer power-of-two
scale S = 2^(14 - ceil(log2 max|dz|)).  Reported: relative Frobenius error of each layer's weight
gradient and of the module's packed gradient with those layers replaced (everything else exact).

    python tools/c8w_estimate.py [--batch 8] [--frames 64]

Test / analysis tooling only (imports the staged reference; never imported by the product).
"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def fp16_round(t, scale=1.0):
    return (t * scale).to(torch.float16).to(torch.float64) / scale


def estimate(batch, frames, verbose=False):
    """{'G': added relative error of the packed Generator gradient, 'D': ..., 'layers': [per-layer rel. err]}"""
    import importlib.util
    # the unmodified reference modules, loaded by PATH (another `mask_cyclegan_vc.model` -- the engine's shim --
    # may already sit in sys.modules of the calling process)
    spec = importlib.util.spec_from_file_location("mcgvc_ref_model_for_c8w_estimate",
                                                  os.path.join(ROOT, "oracle", "_ref", "mask_cyclegan_vc", "model.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    Generator, Discriminator = ref.Generator, ref.Discriminator
    import maskcyclegan_oracle as O
    torch.manual_seed(0)
    G, D = Generator(), Discriminator()
    x, m, _, _ = O.synthetic_batch(batch, frames, seed=11)

    # the C8 layers: Generator ds1/ds2 (conv and gate convs), up1/up2; Discriminator ds1-3
    c8 = {"G": [G.downSample1.convLayer[0], G.downSample1.convLayer_gates[0], G.downSample2.convLayer[0],
                G.downSample2.convLayer_gates[0], G.upSample1[0], G.upSample2[0]],
          "D": [D.downSample1[0], D.downSample2[0], D.downSample3[0]]}
    cap = {}

    def fwd_hook(mod, inp, out):
        cap[mod] = [inp[0].detach(), None]
        out.register_hook(lambda g, mod=mod: cap[mod].__setitem__(1, g.detach()))

    for mods in c8.values():
        for mod in mods:
            mod.register_forward_hook(fwd_hook)
    fake = G(x, m)
    loss = torch.mean((1 - D(fake)) ** 2)
    loss.backward()

    res = {"layers": []}
    for name, net in (("G", G), ("D", D)):
        exact = torch.cat([p.grad.flatten().double() for p in net.parameters() if p.grad is not None])
        err2 = 0.0
        for mod in c8[name]:
            xi, dz = cap[mod]
            S = 2.0 ** (14 - math.ceil(math.log2(float(dz.abs().max()))))
            w64 = torch.nn.grad.conv2d_weight(xi.double(), mod.weight.shape, dz.double(), stride=mod.stride, padding=mod.padding)
            w16 = torch.nn.grad.conv2d_weight(fp16_round(xi), mod.weight.shape, fp16_round(dz, S), stride=mod.stride, padding=mod.padding)
            e = (w16 - w64).norm().item()
            err2 += e * e
            res["layers"].append(e / w64.norm().item())
            if verbose:
                print("%s conv %-22s dW %s: fp16 single pass rel. err %.2e (fp32 autograd vs fp64: %.1e)"
                      % (name, "%dx%d s%d %d->%d" % (mod.kernel_size[0], mod.kernel_size[1], mod.stride[0], mod.in_channels, mod.out_channels),
                         tuple(mod.weight.shape), e / w64.norm().item(), (mod.weight.grad.double() - w64).norm().item() / w64.norm().item()))
        res[name] = math.sqrt(err2) / exact.norm().item()
        if verbose:
            print("%s packed gradient: error added by the fp16 weight-gradient GEMMs = %.2e of its norm (gate 1e-3)" % (name, res[name]))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=64)
    args = ap.parse_args()
    estimate(args.batch, args.frames, verbose=True)


if __name__ == "__main__":
    main()
