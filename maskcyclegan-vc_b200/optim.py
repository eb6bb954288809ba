"""Fused Adam over the engine modules' flat parameter / gradient buffers.

`torch.optim.Adam(g_params, lr, betas=(0.5, 0.999))` at train.py:119-122 walks 220 (generators) or
80 (discriminators) tensors per step; because every engine module already keeps its parameters and
gradients in one flat buffer each, the same update is ONE bandwidth-bound kernel per module here.
Opt-in (train.py builds torch.optim.Adam itself): same defaults and update rule as torch.optim.Adam
without weight decay / amsgrad.  Parameters that never receive a gradient (the Discriminator's
unused downSample4) see g = 0 forever, so their update is exactly zero, as with torch's skip.
"""
import ctypes

import torch

from . import engine


class FusedAdam:
    def __init__(self, modules, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.modules = list(modules)
        self.lr = float(lr)
        self.betas = (float(betas[0]), float(betas[1]))
        self.eps = float(eps)
        self.state = {}
        # torch.optim-compatible view for code that adjusts the learning rate (train.py:139-153)
        self.param_groups = [{"lr": self.lr, "betas": self.betas, "eps": self.eps,
                              "params": [p for m in self.modules for p in m.parameters()]}]

    def zero_grad(self, set_to_none=True):
        for m in self.modules:
            for p in m.parameters():
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        lib = engine.lib()
        lr = float(self.param_groups[0]["lr"])
        for m in self.modules:
            live = m._live_params()
            if not live or live[0].grad is None:
                continue                    # module took no part in the last backward pass
            flat, grad = m._flat, m._flat_grad
            st = self.state.get(id(m))
            if st is None or st["exp_avg"].device != flat.device or st["exp_avg"].numel() != flat.numel():
                st = {"step": 0, "exp_avg": torch.zeros_like(flat), "exp_avg_sq": torch.zeros_like(flat)}
                self.state[id(m)] = st
            st["step"] += 1
            lib.mcgvc_set_device(flat.device.index)
            # the gradient buffer uses the live layout (flat layout minus the never-used downSample4
            # range): one launch per live range; the dead range sees no update, as with torch's skip
            dead_b, dead_n = engine.dead_param_range(m.MODEL)
            ranges = [(0, 0, flat.numel())] if dead_n == 0 else \
                [(0, 0, dead_b), (dead_b + dead_n, dead_b, flat.numel() - dead_b - dead_n)]
            for p_off, g_off, n in ranges:
                if n <= 0:
                    continue
                rc = lib.mcgvc_adam_step(ctypes.c_void_p(flat.data_ptr() + 4 * p_off),
                                         ctypes.c_void_p(grad.data_ptr() + 4 * g_off),
                                         ctypes.c_void_p(st["exp_avg"].data_ptr() + 4 * p_off),
                                         ctypes.c_void_p(st["exp_avg_sq"].data_ptr() + 4 * p_off),
                                         ctypes.c_longlong(n), ctypes.c_float(lr),
                                         ctypes.c_float(self.betas[0]), ctypes.c_float(self.betas[1]),
                                         ctypes.c_float(self.eps), ctypes.c_int(st["step"]),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                if rc != 0:
                    raise engine.EngineError("adam_step failed: " + lib.mcgvc_last_error().decode())
            m._weights_epoch += 1           # the flat buffer changed behind the Parameters' version counters

    def state_dict(self):
        return {"lr": self.param_groups[0]["lr"], "betas": self.betas, "eps": self.eps,
                "state": [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.state.get(id(m), {}).items()}
                          for m in self.modules]}

    def load_state_dict(self, sd):
        self.param_groups[0]["lr"] = sd["lr"]
        self.betas, self.eps = tuple(sd["betas"]), sd["eps"]
        for m, st in zip(self.modules, sd["state"]):
            if st:
                self.state[id(m)] = {k: (v.clone().to(m._flat.device) if torch.is_tensor(v) else v) for k, v in st.items()}
