"""The caller of the hot path: one MaskCycleGAN-VC optimisation step with the semantics of the
reference's training loop body (mask_cyclegan_vc/train.py:186-299), for modules exposing the
reference's Generator / Discriminator call signatures.  bench.py and the GPU tests drive the
engine modules through this; the reference's own train.py can be run unchanged instead via the
PYTHONPATH shim.
"""
import torch


def build_models(generator_cls, discriminator_cls, device, seed=0):
    """Construction order and seeding of train.py:103-110."""
    torch.manual_seed(seed)
    mods = [generator_cls(), generator_cls(), discriminator_cls(), discriminator_cls(),
            discriminator_cls(), discriminator_cls()]
    return [m.to(device) for m in mods]


def build_optimizers(models, g_lr=2e-4, d_lr=1e-4):
    """train.py:113-122."""
    G_A2B, G_B2A, D_A, D_B, D_A2, D_B2 = models
    g_params = list(G_A2B.parameters()) + list(G_B2A.parameters())
    d_params = list(D_A.parameters()) + list(D_B.parameters()) + list(D_A2.parameters()) + \
        list(D_B2.parameters())
    g_opt = torch.optim.Adam(g_params, lr=g_lr, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(d_params, lr=d_lr, betas=(0.5, 0.999))
    return g_opt, d_opt


def train_step(models, g_opt, d_opt, batch, cycle_lambda=10.0, identity_lambda=5.0, fused_losses=False):
    """Returns (g_loss, d_loss) as 0-dim tensors (no host sync here).  fused_losses=True evaluates the
    loss tail with the engine's one-launch-per-term kernels (losses.py) instead of torch ops."""
    if fused_losses:
        from . import losses as fl
    G_A2B, G_B2A, D_A, D_B, D_A2, D_B2 = models
    real_A, mask_A, real_B, mask_B = batch
    # ---- generator phase (train.py:195-242)
    G_A2B.train(); G_B2A.train()
    D_A.eval(); D_B.eval(); D_A2.eval(); D_B2.eval()
    fake_B = G_A2B(real_A, mask_A)
    cycle_A = G_B2A(fake_B, torch.ones_like(fake_B))
    fake_A = G_B2A(real_B, mask_B)
    cycle_B = G_A2B(fake_A, torch.ones_like(fake_A))
    identity_A = G_B2A(real_A, torch.ones_like(real_A))
    identity_B = G_A2B(real_B, torch.ones_like(real_B))
    d_fake_A = D_A(fake_A)
    d_fake_B = D_B(fake_B)
    d_fake_cycle_A = D_A2(cycle_A)
    d_fake_cycle_B = D_B2(cycle_B)
    if fused_losses:
        g_loss = fl.generator_loss(real_A, real_B, cycle_A, cycle_B, identity_A, identity_B, d_fake_A, d_fake_B,
                                   d_fake_cycle_A, d_fake_cycle_B, cycle_lambda, identity_lambda)
    else:
        cycle_loss = torch.mean(torch.abs(real_A - cycle_A)) + torch.mean(torch.abs(real_B - cycle_B))
        identity_loss = torch.mean(torch.abs(real_A - identity_A)) + torch.mean(torch.abs(real_B - identity_B))
        g_loss = torch.mean((1 - d_fake_B) ** 2) + torch.mean((1 - d_fake_A) ** 2) + \
            torch.mean((1 - d_fake_cycle_B) ** 2) + torch.mean((1 - d_fake_cycle_A) ** 2) + \
            cycle_lambda * cycle_loss + identity_lambda * identity_loss
    g_opt.zero_grad()
    d_opt.zero_grad()
    g_loss.backward()
    g_opt.step()
    # ---- discriminator phase (train.py:247-299)
    G_A2B.eval(); G_B2A.eval()
    D_A.train(); D_B.train(); D_A2.train(); D_B2.train()
    d_real_A = D_A(real_A)
    d_real_B = D_B(real_B)
    d_real_A2 = D_A2(real_A)
    d_real_B2 = D_B2(real_B)
    generated_A = G_B2A(real_B, mask_B)
    d_fake_A = D_A(generated_A)
    cycled_B = G_A2B(generated_A, torch.ones_like(generated_A))
    d_cycled_B = D_B2(cycled_B)
    generated_B = G_A2B(real_A, mask_A)
    d_fake_B = D_B(generated_B)
    cycled_A = G_B2A(generated_B, torch.ones_like(generated_B))
    d_cycled_A = D_A2(cycled_A)
    if fused_losses:
        d_loss = fl.discriminator_loss(d_real_A, d_real_B, d_real_A2, d_real_B2, d_fake_A, d_fake_B, d_cycled_A,
                                       d_cycled_B)
    else:
        d_loss_A = (torch.mean((1 - d_real_A) ** 2) + torch.mean((0 - d_fake_A) ** 2)) / 2.0
        d_loss_B = (torch.mean((1 - d_real_B) ** 2) + torch.mean((0 - d_fake_B) ** 2)) / 2.0
        d_loss_A_2nd = (torch.mean((1 - d_real_A2) ** 2) + torch.mean((0 - d_cycled_A) ** 2)) / 2.0
        d_loss_B_2nd = (torch.mean((1 - d_real_B2) ** 2) + torch.mean((0 - d_cycled_B) ** 2)) / 2.0
        d_loss = (d_loss_A + d_loss_B) / 2.0 + (d_loss_A_2nd + d_loss_B_2nd) / 2.0
    g_opt.zero_grad()
    d_opt.zero_grad()
    d_loss.backward()
    d_opt.step()
    return g_loss.detach(), d_loss.detach()


# Algorithmic conv FLOPs (2 x MACs, convolutions only) per 80x64 sample, SURVEY.md 8(d).
G_FWD_FLOPS_T64 = 19676266496.0
D_FWD_FLOPS_T64 = 2277212160.0
STEP_FLOPS_STRICT_T64 = 669.861e9   # everything the reference loop executes, per sample pair
STEP_FLOPS_LEAN_T64 = 504.081e9     # without the gradients train.py discards
