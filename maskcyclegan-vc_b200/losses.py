"""Fused loss tail: the generator and discriminator losses of
mask_cyclegan_vc/train.py:219-237 and :276-294 as one reduction launch per term (all terms of a phase
accumulate, already weighted, into one device scalar) and one elementwise launch per term backward,
instead of ~4 aten kernels per term.  Opt-in: the unchanged reference train.py keeps its torch ops.
"""
import ctypes

import torch
from torch.autograd import Function

from . import engine

L1, LSGAN = 0, 1


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _WeightedSum(Function):
    """total = sum_k weight_k * term_k(a_k[, b_k]); terms = [(kind, target, weight, has_b)], tensors flat list."""

    @staticmethod
    def forward(ctx, terms, *tensors):
        lib = engine.lib()
        dev = tensors[0].device
        lib.mcgvc_set_device(dev.index or 0)
        total = torch.zeros((), dtype=torch.float32, device=dev)
        pairs, i = [], 0
        for kind, target, weight, has_b in terms:
            a = tensors[i].contiguous()
            b = tensors[i + 1].contiguous() if has_b else None
            i += 2 if has_b else 1
            if not a.is_cuda or a.dtype != torch.float32 or (b is not None and b.shape != a.shape):
                raise engine.EngineError("loss terms take float32 CUDA tensors of matching shapes")
            rc = lib.mcgvc_loss_term(_p(a), _p(b), ctypes.c_longlong(a.numel()), kind, ctypes.c_float(target),
                                     ctypes.c_float(weight), _p(total), _stream())
            if rc != 0:
                raise engine.EngineError("loss_term failed: %s" % lib.mcgvc_last_error().decode())
            pairs.append((a, b))
        ctx.terms = terms
        ctx.has_b = [t[3] for t in terms]
        ctx.save_for_backward(*[t for pr in pairs for t in pr if t is not None])
        return total

    @staticmethod
    def backward(ctx, gtotal):
        lib = engine.lib()
        saved = list(ctx.saved_tensors)
        gtotal = gtotal.contiguous().float()
        grads, i, k = [None], 0, 0
        for (kind, target, weight, has_b) in ctx.terms:
            a = saved[i]
            b = saved[i + 1] if has_b else None
            i += 2 if has_b else 1
            ga = None
            if ctx.needs_input_grad[1 + k]:
                ga = torch.empty_like(a)
                rc = lib.mcgvc_loss_term_grad(_p(a), _p(b), ctypes.c_longlong(a.numel()), kind, ctypes.c_float(target),
                                              ctypes.c_float(weight), _p(gtotal), _p(ga), _stream())
                if rc != 0:
                    raise engine.EngineError("loss_term_grad failed: %s" % lib.mcgvc_last_error().decode())
            grads.append(ga)
            k += 1
            if has_b:
                grads.append(None)     # b is data in every term of the reference's losses
                k += 1
        return tuple(grads)


def weighted_loss(terms):
    """terms: list of (kind, a, b_or_None, target, weight).  Returns the 0-dim total (differentiable in each a)."""
    spec, tensors = [], []
    for kind, a, b, target, weight in terms:
        spec.append((kind, float(target), float(weight), b is not None))
        tensors.append(a)
        if b is not None:
            tensors.append(b.detach())
    return _WeightedSum.apply(tuple(spec), *tensors)


def generator_loss(real_A, real_B, cycle_A, cycle_B, identity_A, identity_B, d_fake_A, d_fake_B, d_fake_cycle_A,
                   d_fake_cycle_B, cycle_lambda=10.0, identity_lambda=5.0):
    """g_loss of train.py:219-237."""
    return weighted_loss([
        (LSGAN, d_fake_B, None, 1.0, 1.0), (LSGAN, d_fake_A, None, 1.0, 1.0),
        (LSGAN, d_fake_cycle_B, None, 1.0, 1.0), (LSGAN, d_fake_cycle_A, None, 1.0, 1.0),
        (L1, cycle_A, real_A, 0.0, cycle_lambda), (L1, cycle_B, real_B, 0.0, cycle_lambda),
        (L1, identity_A, real_A, 0.0, identity_lambda), (L1, identity_B, real_B, 0.0, identity_lambda)])


def discriminator_loss(d_real_A, d_real_B, d_real_A2, d_real_B2, d_fake_A, d_fake_B, d_cycled_A, d_cycled_B):
    """d_loss of train.py:276-294: every one of the eight terms ends up with weight 1/4."""
    return weighted_loss([(LSGAN, d, None, 1.0, 0.25) for d in (d_real_A, d_real_B, d_real_A2, d_real_B2)] +
                         [(LSGAN, d, None, 0.0, 0.25) for d in (d_fake_A, d_fake_B, d_cycled_A, d_cycled_B)])
