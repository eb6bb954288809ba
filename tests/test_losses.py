"""Fused loss tail against the reference's loss expressions
(mask_cyclegan_vc/train.py:219-237 and :276-294) evaluated with torch on the CPU."""
import pytest
import torch

import maskcyclegan_oracle as O

pytestmark = pytest.mark.gpu


def _ref_g_loss(real_A, real_B, cycle_A, cycle_B, identity_A, identity_B, dfa, dfb, dca, dcb, lc=10.0, li=5.0):
    cycle = torch.mean(torch.abs(real_A - cycle_A)) + torch.mean(torch.abs(real_B - cycle_B))          # :219-220
    ident = torch.mean(torch.abs(real_A - identity_A)) + torch.mean(torch.abs(real_B - identity_B))    # :223-224
    return torch.mean((1 - dfb) ** 2) + torch.mean((1 - dfa) ** 2) + torch.mean((1 - dcb) ** 2) + \
        torch.mean((1 - dca) ** 2) + lc * cycle + li * ident                                          # :227-237


def _ref_d_loss(ra, rb, ra2, rb2, fa, fb, ca, cb):
    d_a = (torch.mean((1 - ra) ** 2) + torch.mean((0 - fa) ** 2)) / 2.0                                 # :276-278
    d_b = (torch.mean((1 - rb) ** 2) + torch.mean((0 - fb) ** 2)) / 2.0                                 # :280-282
    d_a2 = (torch.mean((1 - ra2) ** 2) + torch.mean((0 - ca) ** 2)) / 2.0                               # :285-290
    d_b2 = (torch.mean((1 - rb2) ** 2) + torch.mean((0 - cb) ** 2)) / 2.0
    return (d_a + d_b) / 2.0 + (d_a2 + d_b2) / 2.0                                                      # :293-294


def test_generator_and_discriminator_losses_match_the_reference_expressions(pkg):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = torch.Generator().manual_seed(0)
    mel = [torch.randn(4, 80, 64, generator=g) for _ in range(6)]
    mel[2] = mel[0].clone()
    mel[2][0, :5] += 0.25                       # exact ties elsewhere: sign(0) = 0 in the L1 gradient
    dd = [torch.rand(4, 1, 10, 8, generator=g) for _ in range(8)]
    # generator loss: gradients w.r.t. the four G outputs and the four D outputs
    cpu = [t.clone().requires_grad_(i >= 2) for i, t in enumerate(mel)] + [t.clone().requires_grad_(True) for t in dd[:4]]
    ref = _ref_g_loss(*cpu)
    ref.backward()
    dev = [t.cuda().requires_grad_(i >= 2) for i, t in enumerate(mel)] + [t.cuda().requires_grad_(True) for t in dd[:4]]
    got = pkg.losses.generator_loss(*dev)
    (got * 1.5).backward()                      # non-trivial upstream gradient
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item())
    for c, d in zip(cpu[2:], dev[2:]):
        assert torch.allclose(d.grad.cpu(), 1.5 * c.grad, rtol=1e-5, atol=1e-9)
    assert dev[0].grad is None and dev[1].grad is None
    # discriminator loss
    cpu = [t.clone().requires_grad_(True) for t in dd]
    ref = _ref_d_loss(*cpu)
    ref.backward()
    dev = [t.cuda().requires_grad_(True) for t in dd]
    got = pkg.losses.discriminator_loss(*dev)
    got.backward()
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item())
    for c, d in zip(cpu, dev):
        assert torch.allclose(d.grad.cpu(), c.grad, rtol=1e-5, atol=1e-9)
    with pytest.raises(pkg.engine.EngineError):
        pkg.losses.weighted_loss([(pkg.losses.L1, mel[0], mel[1], 0.0, 1.0)])   # CPU tensors: no fallback


def test_train_step_with_fused_losses_follows_the_torch_tail(pkg):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from maskcyclegan_vc_b200 import trainstep as ts
    e = pkg.engine
    e.set_precision(e.PRECISION_PARITY)
    out = []
    for fused in (False, True):
        models = ts.build_models(pkg.Generator, pkg.Discriminator, torch.device("cuda"), seed=0)
        g_opt, d_opt = ts.build_optimizers(models)
        losses = []
        for step in range(2):
            batch = [t.cuda() for t in O.synthetic_batch(2, 64, seed=900 + step)]
            gl, dl = ts.train_step(models, g_opt, d_opt, batch, fused_losses=fused)
            losses.append((gl.item(), dl.item()))
        out.append(losses)
    for (ga, da), (gb, db) in zip(*out):
        assert abs(ga - gb) <= 1e-4 * abs(ga) and abs(da - db) <= 1e-4 * abs(da) + 1e-6
