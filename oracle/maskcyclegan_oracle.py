"""CPU oracle for the MaskCycleGAN-VC Generator / Discriminator hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; it is used by
tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.

It restates, in plain functional PyTorch on CPU (fp32 or fp64), what the reference computes in
/root/reference/mask_cyclegan_vc/model.py and the train step in
/root/reference/mask_cyclegan_vc/train.py:186-299.  The arithmetic itself lives in PyTorch (the
reference pins pytorch=1.7.1, environment.yml:121; conv / instance_norm / pixel_shuffle / sigmoid
semantics are unchanged in the torch 2.11 used here).

Parity pinning: the reference has NO golden vectors or known-answer tests (only the shape check at
model.py:352-371).  This oracle is therefore pinned against outputs of the reference itself:
oracle/make_golden.py imports the real model.py from /root/reference in the build container and
commits inputs/outputs under tests/golden/; tests/test_oracle.py checks this restatement against
those fixtures (and against a live import of the reference when /root/reference is present).
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

G_PARAM_COUNT = 24537729
D_PARAM_COUNT = 16691713


# ------------------------------------------------------------------------------------------------
# Parameter construction: replays the reference's construction order so that torch.manual_seed(s)
# followed by build_*_state() yields exactly the tensors `Generator()` / `Discriminator()` would
# hold (model.py:110-211 and :287-327).  Uses throw-away nn modules so torch's own default init
# (kaiming_uniform_(a=sqrt(5)) + bias uniform) consumes the RNG in the same order.
def _conv2d(sd, name, cin, cout, k, stride, pad):
    m = nn.Conv2d(cin, cout, k, stride, pad)
    sd[name + ".weight"] = m.weight.detach().clone()
    sd[name + ".bias"] = m.bias.detach().clone()


def _conv1d(sd, name, cin, cout, k, pad):
    m = nn.Conv1d(cin, cout, k, 1, pad)
    sd[name + ".weight"] = m.weight.detach().clone()
    sd[name + ".bias"] = m.bias.detach().clone()


def _inorm(sd, name, c):
    sd[name + ".weight"] = torch.ones(c)
    sd[name + ".bias"] = torch.zeros(c)


def build_generator_state(residual_in_channels=256, cx=80):
    """Ordered state of the reference Generator (model.py:110-211), 110 unique tensors."""
    r = residual_in_channels
    flat = (cx // 4) * r
    sd = OrderedDict()
    _conv2d(sd, "conv1", 2, r // 2, (5, 15), 1, (2, 7))
    _conv2d(sd, "conv1_gates", 2, r // 2, (5, 15), 1, (2, 7))
    for name, cin in (("downSample1", r // 2), ("downSample2", r)):
        _conv2d(sd, name + ".convLayer.0", cin, r, 5, 2, 2)
        _inorm(sd, name + ".convLayer.1", r)
        _conv2d(sd, name + ".convLayer_gates.0", cin, r, 5, 2, 2)
        _inorm(sd, name + ".convLayer_gates.1", r)
    _conv1d(sd, "conv2dto1dLayer", flat, r, 1, 0)
    _inorm(sd, "conv2dto1dLayer_tfan", r)
    for i in range(1, 7):
        p = "residualLayer%d" % i
        _conv1d(sd, p + ".conv1d_layer.0", r, 2 * r, 3, 1)
        _inorm(sd, p + ".conv1d_layer.1", 2 * r)
        _conv1d(sd, p + ".conv_layer_gates.0", r, 2 * r, 3, 1)
        _inorm(sd, p + ".conv_layer_gates.1", 2 * r)
        _conv1d(sd, p + ".conv1d_out_layer.0", 2 * r, r, 3, 1)
        _inorm(sd, p + ".conv1d_out_layer.1", r)
    _conv1d(sd, "conv1dto2dLayer", r, flat, 1, 0)
    _inorm(sd, "conv1dto2dLayer_tfan", flat)
    _conv2d(sd, "upSample1.0", r, 4 * r, 5, 1, 2)
    _inorm(sd, "upSample1.2", r)
    _conv2d(sd, "upSample2.0", r, 2 * r, 5, 1, 2)
    _inorm(sd, "upSample2.2", r // 2)
    _conv2d(sd, "lastConvLayer", r // 2, 1, (5, 15), 1, (2, 7))
    return sd


def build_discriminator_state(residual_in_channels=256):
    """Ordered state of the reference Discriminator (model.py:287-327), 20 tensors incl. the unused
    downSample4 (constructed at :316-320, never called in forward :340-349)."""
    r = residual_in_channels
    sd = OrderedDict()
    _conv2d(sd, "convLayer1.0", 1, r // 2, (3, 3), 1, 1)
    for name, cin, cout in (("downSample1", r // 2, r), ("downSample2", r, 2 * r),
                            ("downSample3", 2 * r, 4 * r)):
        _conv2d(sd, name + ".0", cin, cout, 3, 2, 1)
        _inorm(sd, name + ".1", cout)
    _conv2d(sd, "downSample4.0", 4 * r, 4 * r, (1, 10), 1, (0, 2))
    _inorm(sd, "downSample4.1", 4 * r)
    _conv2d(sd, "outputConvLayer.0", 4 * r, 1, (1, 3), 1, (0, 1))
    return sd


def reference_state_dict_keys_generator(sd):
    """The reference's state_dict() has 114 keys: `convLayer.*` aliases `upSample2.*` because
    Generator.upsample() assigns self.convLayer (model.py:227); order follows module registration."""
    out = OrderedDict()
    for k, v in sd.items():
        if k.startswith("upSample1."):
            break
        out[k] = v
    for k, v in sd.items():
        if k.startswith("upSample2."):
            out["convLayer." + k[len("upSample2."):]] = v
    for k, v in sd.items():
        if k.startswith("upSample1.") or k.startswith("upSample2.") or k.startswith("lastConvLayer"):
            out[k] = v
    return out


# ------------------------------------------------------------------------------------------------
def _inst_norm(x, w, b):
    # nn.InstanceNorm{1,2}d(affine=True, track_running_stats=False): biased variance, eps=1e-5
    return F.instance_norm(x, None, None, w, b, True, 0.1, 1e-5)


def _swish(x):  # model.py:12-21 "GLU" is x * sigmoid(x)
    return x * torch.sigmoid(x)


def generator_forward(sd, x, mask, taps=None):
    """model.py:239-280.  `sd` maps reference parameter names to tensors; `taps`, when a dict, is
    filled with named intermediates (NCHW / NCL as the reference holds them)."""
    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    h = torch.stack((x * mask, mask), dim=1)                                    # :241
    conv1 = F.conv2d(h, sd["conv1.weight"], sd["conv1.bias"], 1, (2, 7)) * \
        torch.sigmoid(F.conv2d(h, sd["conv1_gates.weight"], sd["conv1_gates.bias"], 1, (2, 7)))  # :242
    rec("conv1", conv1)
    h = conv1
    for name in ("downSample1", "downSample2"):                                 # :245-246, :101-103
        a = _inst_norm(F.conv2d(h, sd[name + ".convLayer.0.weight"], sd[name + ".convLayer.0.bias"], 2, 2),
                       sd[name + ".convLayer.1.weight"], sd[name + ".convLayer.1.bias"])
        g = _inst_norm(F.conv2d(h, sd[name + ".convLayer_gates.0.weight"], sd[name + ".convLayer_gates.0.bias"], 2, 2),
                       sd[name + ".convLayer_gates.1.weight"], sd[name + ".convLayer_gates.1.bias"])
        h = rec(name, a * torch.sigmoid(g))
    flat = sd["conv2dto1dLayer.weight"].shape[1]
    h = h.view(h.size(0), flat, 1, -1).squeeze(2)                               # :249-251
    h = F.conv1d(h, sd["conv2dto1dLayer.weight"], sd["conv2dto1dLayer.bias"])   # :254
    h = rec("conv2dto1d", _inst_norm(h, sd["conv2dto1dLayer_tfan.weight"], sd["conv2dto1dLayer_tfan.bias"]))
    for i in range(1, 7):                                                       # :258-263, :71-76
        p = "residualLayer%d" % i
        a = _inst_norm(F.conv1d(h, sd[p + ".conv1d_layer.0.weight"], sd[p + ".conv1d_layer.0.bias"], 1, 1),
                       sd[p + ".conv1d_layer.1.weight"], sd[p + ".conv1d_layer.1.bias"])
        g = _inst_norm(F.conv1d(h, sd[p + ".conv_layer_gates.0.weight"], sd[p + ".conv_layer_gates.0.bias"], 1, 1),
                       sd[p + ".conv_layer_gates.1.weight"], sd[p + ".conv_layer_gates.1.bias"])
        glu = a * torch.sigmoid(g)
        o = _inst_norm(F.conv1d(glu, sd[p + ".conv1d_out_layer.0.weight"], sd[p + ".conv1d_out_layer.0.bias"], 1, 1),
                       sd[p + ".conv1d_out_layer.1.weight"], sd[p + ".conv1d_out_layer.1.bias"])
        h = rec(p, h + o)
    h = F.conv1d(h, sd["conv1dto2dLayer.weight"], sd["conv1dto2dLayer.bias"])   # :266
    h = _inst_norm(h, sd["conv1dto2dLayer_tfan.weight"], sd["conv1dto2dLayer_tfan.bias"])
    h = rec("conv1dto2d", h.unsqueeze(2).view(h.size(0), 256, 20, -1))          # :270-271
    for name in ("upSample1", "upSample2"):                                     # :274-275, :226-237
        h = F.conv2d(h, sd[name + ".0.weight"], sd[name + ".0.bias"], 1, 2)
        h = F.pixel_shuffle(h, 2)
        h = _inst_norm(h, sd[name + ".2.weight"], sd[name + ".2.bias"])
        h = rec(name, _swish(h))
    out = F.conv2d(h, sd["lastConvLayer.weight"], sd["lastConvLayer.bias"], 1, (2, 7))  # :278
    return out.squeeze(1)                                                       # :279


def discriminator_forward(sd, x, taps=None):
    """model.py:340-349 (downSample4 is never applied)."""
    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    h = x.unsqueeze(1)                                                          # :343
    h = rec("convLayer1", _swish(F.conv2d(h, sd["convLayer1.0.weight"], sd["convLayer1.0.bias"], 1, 1)))
    for name in ("downSample1", "downSample2", "downSample3"):                  # :345-347
        h = F.conv2d(h, sd[name + ".0.weight"], sd[name + ".0.bias"], 2, 1)
        h = rec(name, _swish(_inst_norm(h, sd[name + ".1.weight"], sd[name + ".1.bias"])))
    h = F.conv2d(h, sd["outputConvLayer.0.weight"], sd["outputConvLayer.0.bias"], 1, (0, 1))
    return torch.sigmoid(h)                                                     # :348


# ------------------------------------------------------------------------------------------------
def make_fif_mask(shape, max_mask_len, gen):
    """Filling-in-frames mask, dataset/vc_dataset.py:51-55 semantics on a batch (B, 80, T):
    size ~ U{0..max_mask_len-1}, start ~ U{0..T-size-1}; masked frames are 0 across all bins."""
    B, C, T = shape
    m = torch.ones(shape)
    for b in range(B):
        size = int(torch.randint(0, max_mask_len, (1,), generator=gen))
        start = int(torch.randint(0, T - size, (1,), generator=gen)) if T - size > 0 else 0
        m[b, :, start:start + size] = 0.0
    return m


def synthetic_batch(B, T, seed, max_mask_len=25):
    """real_A/B ~ N(0,1), FIF masks; deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    real_A = torch.randn(B, 80, T, generator=g)
    real_B = torch.randn(B, 80, T, generator=g)
    mask_A = make_fif_mask((B, 80, T), max_mask_len, g)
    mask_B = make_fif_mask((B, 80, T), max_mask_len, g)
    return real_A, mask_A, real_B, mask_B


def crop_and_mask(dataset, sel, n_frames=64):
    """numpy restatement of the per-sample crop + frame-in-fill mask of the reference's training
    dataset (dataset/vc_dataset.py:44-56) for given draws: sel = int[4][B] (utterance, crop start,
    mask start, mask size).  Returns (x, mask), each (B, 80, n_frames) float32."""
    import numpy as np
    xs, ms = [], []
    for u, start, mstart, msize in zip(*[list(r) for r in sel]):
        data = np.asarray(dataset[u])
        crop = data[:, start:start + n_frames]                 # vc_dataset.py:48-49,55
        mask = np.ones_like(crop)                              # vc_dataset.py:53
        mask[:, mstart:mstart + msize] = 0.                    # vc_dataset.py:54
        xs.append(crop)
        ms.append(mask)
    return np.array(xs, dtype=np.float32), np.array(ms, dtype=np.float32)


def train_step(G_A2B, G_B2A, D_A, D_B, D_A2, D_B2, g_opt, d_opt, batch,
               cycle_lambda=10.0, identity_lambda=5.0):
    """One optimisation step with the semantics of train.py:186-299 for any modules that expose
    the reference's Generator/Discriminator call signatures.  Returns (g_loss, d_loss) floats."""
    real_A, mask_A, real_B, mask_B = batch
    for g in (G_A2B, G_B2A):
        g.train()
    for d in (D_A, D_B, D_A2, D_B2):
        d.eval()
    fake_B = G_A2B(real_A, mask_A)                                              # :203
    cycle_A = G_B2A(fake_B, torch.ones_like(fake_B))
    fake_A = G_B2A(real_B, mask_B)
    cycle_B = G_A2B(fake_A, torch.ones_like(fake_A))
    identity_A = G_B2A(real_A, torch.ones_like(real_A))
    identity_B = G_A2B(real_B, torch.ones_like(real_B))                         # :210
    d_fake_A = D_A(fake_A)
    d_fake_B = D_B(fake_B)
    d_fake_cycle_A = D_A2(cycle_A)
    d_fake_cycle_B = D_B2(cycle_B)                                              # :216
    cycle_loss = torch.mean(torch.abs(real_A - cycle_A)) + torch.mean(torch.abs(real_B - cycle_B))
    identity_loss = torch.mean(torch.abs(real_A - identity_A)) + torch.mean(torch.abs(real_B - identity_B))
    g_loss = torch.mean((1 - d_fake_B) ** 2) + torch.mean((1 - d_fake_A) ** 2) + \
        torch.mean((1 - d_fake_cycle_B) ** 2) + torch.mean((1 - d_fake_cycle_A) ** 2) + \
        cycle_lambda * cycle_loss + identity_lambda * identity_loss            # :235-237
    g_opt.zero_grad()
    d_opt.zero_grad()
    g_loss.backward()
    g_opt.step()                                                                # :240-242

    for g in (G_A2B, G_B2A):
        g.eval()
    for d in (D_A, D_B, D_A2, D_B2):
        d.train()
    d_real_A = D_A(real_A)                                                      # :255
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
    d_cycled_A = D_A2(cycled_A)                                                 # :273
    d_loss_A = (torch.mean((1 - d_real_A) ** 2) + torch.mean((0 - d_fake_A) ** 2)) / 2.0
    d_loss_B = (torch.mean((1 - d_real_B) ** 2) + torch.mean((0 - d_fake_B) ** 2)) / 2.0
    d_loss_A_2nd = (torch.mean((1 - d_real_A2) ** 2) + torch.mean((0 - d_cycled_A) ** 2)) / 2.0
    d_loss_B_2nd = (torch.mean((1 - d_real_B2) ** 2) + torch.mean((0 - d_cycled_B) ** 2)) / 2.0
    d_loss = (d_loss_A + d_loss_B) / 2.0 + (d_loss_A_2nd + d_loss_B_2nd) / 2.0  # :293-294
    g_opt.zero_grad()
    d_opt.zero_grad()
    d_loss.backward()
    d_opt.step()                                                                # :297-299
    return float(g_loss.item()), float(d_loss.item())


class OracleGenerator(nn.Module):
    """nn.Module wrapper over generator_forward with reference-ordered parameters (CPU baseline)."""

    def __init__(self):
        super().__init__()
        sd = build_generator_state()
        self._names = list(sd.keys())
        self.params = nn.ParameterList([nn.Parameter(v) for v in sd.values()])

    def sd(self):
        return {n: p for n, p in zip(self._names, self.params)}

    def forward(self, x, mask):
        return generator_forward(self.sd(), x, mask)


class OracleDiscriminator(nn.Module):
    def __init__(self):
        super().__init__()
        sd = build_discriminator_state()
        self._names = list(sd.keys())
        self.params = nn.ParameterList([nn.Parameter(v) for v in sd.values()])

    def sd(self):
        return {n: p for n, p in zip(self._names, self.params)}

    def forward(self, x):
        return discriminator_forward(self.sd(), x)
