"""ctypes binding of libmcgvc.so (C ABI in include/mcgvc.h).

Raw device pointers and a cudaStream_t go down; nothing here falls back to PyTorch arithmetic.  If
the shared library is missing or the tensors are not CUDA tensors, calls fail loudly.
"""
import ctypes
import os

import torch

GENERATOR = 0
DISCRIMINATOR = 1
BACKEND_TCGEN05 = 0
BACKEND_SIMT = 1
PRECISION_PARITY = 3   # split-bf16 x3 (meets the 1e-3 parity gate; 3 MMA units per MAC)
PRECISION_MIXED = 2    # forward split-bf16 x3, backward single bf16 pass
PRECISION_FAST = 1     # single bf16 pass
PRECISION_C8 = 4       # fp16 main pass + two e4m3 correction passes in every GEMM of the C8 layers (2 MMA units per MAC); meets the 1e-3 gate
PRECISION_C8H = 5      # forward as C8; backward GEMMs of the C8 layers: one fp16 pass (TF32-class gradients)
PRECISION_C8W = 6      # DEFAULT: forward and data gradients as C8; weight-gradient GEMMs of the C8 layers: one fp16 pass (same 1e-3 gate)

# MCGVC_LIBRARY: another build of the same library (A/B timing of two builds in one gpurun call)
_LIB_PATH = os.environ.get("MCGVC_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmcgvc.so")
_lib = None

_c_ll = ctypes.c_longlong
_c_vp = ctypes.c_void_p
_c_int = ctypes.c_int


class EngineError(RuntimeError):
    pass


def lib():
    """Load libmcgvc.so once.  Raises if it has not been built (see __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise EngineError(
            "libmcgvc.so not found at %s: build it with `make -C maskcyclegan-vc_b200/csrc` "
            "(or __graft_entry__.build()); there is no fallback path" % _LIB_PATH)
    l = ctypes.CDLL(_LIB_PATH)
    l.mcgvc_last_error.restype = ctypes.c_char_p
    for name in ("mcgvc_param_count", "mcgvc_packed_bytes", "mcgvc_grad_blob_floats"):
        getattr(l, name).restype = _c_ll
        getattr(l, name).argtypes = [_c_int]
    for name in ("mcgvc_saved_bytes", "mcgvc_fwd_workspace_bytes", "mcgvc_bwd_workspace_bytes"):
        getattr(l, name).restype = _c_ll
        getattr(l, name).argtypes = [_c_int, _c_int, _c_int]
    l.mcgvc_pack_weights.argtypes = [_c_int, _c_vp, _c_vp, _c_vp]
    l.mcgvc_unpack_grads.argtypes = [_c_int, _c_vp, _c_vp, _c_vp]
    l.mcgvc_unpack_grads_live.argtypes = [_c_int, _c_vp, _c_vp, ctypes.c_float, _c_vp]
    l.mcgvc_dead_param_range.argtypes = [_c_int, ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll)]
    l.mcgvc_generator_forward.argtypes = [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp]
    l.mcgvc_generator_backward.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp,
                                           _c_int, _c_vp, _c_vp]
    l.mcgvc_discriminator_forward.argtypes = [_c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp]
    l.mcgvc_discriminator_backward.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp,
                                               _c_int, _c_vp, _c_vp]
    l.mcgvc_adam_step.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp, _c_ll, ctypes.c_float, ctypes.c_float,
                                  ctypes.c_float, ctypes.c_float, _c_int, _c_vp]
    l.mcgvc_saved_layout.argtypes = [_c_int, _c_int, _c_int, _c_int, ctypes.c_char_p, _c_int,
                                     ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll)]
    _lib = l
    return l


def _check(rc, what):
    if rc != 0:
        raise EngineError("%s failed: %s" % (what, lib().mcgvc_last_error().decode()))


def _ptr(t):
    return _c_vp(t.data_ptr()) if t is not None else _c_vp(0)


def _stream():
    return _c_vp(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise EngineError("%s must be a CUDA tensor: the engine has no CPU path" % name)
    if t.dtype != torch.float32:
        raise EngineError("%s must be float32 (got %s)" % (name, t.dtype))


def set_backend(backend):
    _check(lib().mcgvc_set_backend(backend), "set_backend")


def set_precision(mode):
    _check(lib().mcgvc_set_precision(mode), "set_precision")


def set_overlap(on):
    lib().mcgvc_set_overlap(1 if on else 0)


def set_graphs(on):
    lib().mcgvc_set_graphs(1 if on else 0)


def graph_stats():
    c, r = _c_ll(0), _c_ll(0)
    lib().mcgvc_graph_stats(ctypes.byref(c), ctypes.byref(r))
    return {"captures": c.value, "replays": r.value}


def get_precision():
    return lib().mcgvc_get_precision()


def pack_class(precision=None):
    """Modes that share one packed-weight layout: split-bf16 planes (parity / mixed / fast) or the
    fp16 + 2 x e4m3 planes (C8 / C8H)."""
    p = get_precision() if precision is None else precision
    return 1 if p in (PRECISION_C8, PRECISION_C8H, PRECISION_C8W) else 0


def param_count(model):
    return lib().mcgvc_param_count(model)


def packed_bytes(model):
    return lib().mcgvc_packed_bytes(model)


def grad_blob_floats(model):
    return lib().mcgvc_grad_blob_floats(model)


def generator_out_frames(T):
    return lib().mcgvc_generator_out_frames(T)


def discriminator_out_frames(T):
    return lib().mcgvc_discriminator_out_frames(T)


def _bytes(n, device):
    return torch.empty(int(n), dtype=torch.uint8, device=device)


def pack_weights(model, flat_params):
    _require_cuda(flat_params, "parameters")
    l = lib()
    l.mcgvc_set_device(flat_params.device.index)
    packed = _bytes(l.mcgvc_packed_bytes(model), flat_params.device)
    _check(l.mcgvc_pack_weights(model, _ptr(flat_params), _ptr(packed), _stream()), "pack_weights")
    return packed


def unpack_grads(model, grad_blob, flat_grad):
    """flat_grad (full reference-order layout) += engine-layout gradient blob."""
    l = lib()
    l.mcgvc_set_device(flat_grad.device.index)
    _check(l.mcgvc_unpack_grads(model, _ptr(grad_blob), _ptr(flat_grad), _stream()), "unpack_grads")


def dead_param_range(model):
    """(begin, len) in floats of the parameters that never receive a gradient (D: downSample4)."""
    b, n = _c_ll(0), _c_ll(0)
    _check(lib().mcgvc_dead_param_range(model, ctypes.byref(b), ctypes.byref(n)), "dead_param_range")
    return b.value, n.value


def live_grad_count(model):
    return param_count(model) - dead_param_range(model)[1]


def unpack_grads_live(model, grad_blob, live_grad, scale=1.0):
    """live_grad (flat layout minus the dead range) += scale * engine-layout gradient blob."""
    l = lib()
    l.mcgvc_set_device(live_grad.device.index)
    _check(l.mcgvc_unpack_grads_live(model, _ptr(grad_blob), _ptr(live_grad), ctypes.c_float(scale), _stream()),
           "unpack_grads_live")


def generator_forward(packed, x, mask):
    """x, mask: (B, 80, T) fp32 CUDA.  Returns (out (B, 80, T'), saved blob)."""
    _require_cuda(x, "x")
    _require_cuda(mask, "mask")
    if x.dim() != 3 or x.shape[1] != 80 or mask.shape != x.shape:
        raise EngineError("Generator expects x and mask of shape (B, 80, T); got %s and %s"
                          % (tuple(x.shape), tuple(mask.shape)))
    l = lib()
    B, _, T = x.shape
    l.mcgvc_set_device(x.device.index)
    out = torch.empty(B, 80, l.mcgvc_generator_out_frames(T), dtype=torch.float32, device=x.device)
    saved = _bytes(l.mcgvc_saved_bytes(GENERATOR, B, T), x.device)
    ws = _bytes(l.mcgvc_fwd_workspace_bytes(GENERATOR, B, T), x.device)
    _check(l.mcgvc_generator_forward(_ptr(packed), _ptr(x), _ptr(mask), B, T, _ptr(out), _ptr(saved),
                                     _ptr(ws), _stream()), "generator_forward")
    return out, saved


def generator_backward(packed, saved, mask, dout, B, T, need_dx, grad_blob, need_wgrad):
    l = lib()
    l.mcgvc_set_device(dout.device.index)
    dx = torch.empty(B, 80, T, dtype=torch.float32, device=dout.device) if need_dx else None
    ws = _bytes(l.mcgvc_bwd_workspace_bytes(GENERATOR, B, T), dout.device)
    _check(l.mcgvc_generator_backward(_ptr(packed), _ptr(saved), _ptr(mask), _ptr(dout), B, T, _ptr(dx),
                                      _ptr(grad_blob), 1 if need_wgrad else 0, _ptr(ws), _stream()),
           "generator_backward")
    return dx


def discriminator_forward(packed, x):
    """x: (B, 80, T) fp32 CUDA.  Returns (out (B, 1, 10, ceil(T/8)), saved blob)."""
    _require_cuda(x, "x")
    if x.dim() != 3 or x.shape[1] != 80:
        raise EngineError("Discriminator expects x of shape (B, 80, T); got %s" % (tuple(x.shape),))
    l = lib()
    B, _, T = x.shape
    l.mcgvc_set_device(x.device.index)
    out = torch.empty(B, 1, 10, l.mcgvc_discriminator_out_frames(T), dtype=torch.float32, device=x.device)
    saved = _bytes(l.mcgvc_saved_bytes(DISCRIMINATOR, B, T), x.device)
    ws = _bytes(l.mcgvc_fwd_workspace_bytes(DISCRIMINATOR, B, T), x.device)
    _check(l.mcgvc_discriminator_forward(_ptr(packed), _ptr(x), B, T, _ptr(out), _ptr(saved), _ptr(ws),
                                         _stream()), "discriminator_forward")
    return out, saved


def discriminator_backward(packed, saved, out, dout, B, T, need_dx, grad_blob, need_wgrad):
    l = lib()
    l.mcgvc_set_device(dout.device.index)
    dx = torch.empty(B, 80, T, dtype=torch.float32, device=dout.device) if need_dx else None
    ws = _bytes(l.mcgvc_bwd_workspace_bytes(DISCRIMINATOR, B, T), dout.device)
    _check(l.mcgvc_discriminator_backward(_ptr(packed), _ptr(saved), _ptr(out), _ptr(dout), B, T, _ptr(dx),
                                          _ptr(grad_blob), 1 if need_wgrad else 0, _ptr(ws), _stream()),
           "discriminator_backward")
    return dx


def saved_layout(model, B, T):
    """[(name, byte offset, bytes)] of the named tensors inside a saved blob (tests only)."""
    l = lib()
    out = []
    i = 0
    while True:
        name = ctypes.create_string_buffer(64)
        off = _c_ll(0)
        nb = _c_ll(0)
        if l.mcgvc_saved_layout(model, B, T, i, name, 64, ctypes.byref(off), ctypes.byref(nb)) != 0:
            break
        out.append((name.value.decode(), off.value, nb.value))
        i += 1
    return out


def launch_count():
    """Number of kernels libmcgvc.so has launched in this process so far."""
    l = lib()
    l.mcgvc_launch_count.restype = _c_ll
    return l.mcgvc_launch_count()


def profile_enable(on):
    lib().mcgvc_profile_enable(1 if on else 0)


def profile_collect():
    """{'conv': {ms, flops, launches}, 'wgrad': {...}} accumulated since the last collect."""
    buf = (ctypes.c_double * 6)()
    lib().mcgvc_profile_collect(buf)
    return {"conv": {"ms": buf[0], "flops": buf[1], "launches": int(buf[2])},
            "wgrad": {"ms": buf[3], "flops": buf[4], "launches": int(buf[5])}}


def profile_collect_kinds():
    """[{'kernel', 'ms', 'flops', 'launches'}] per tensor-core kernel family since the last collect."""
    l = lib()
    l.mcgvc_profile_kind_name.restype = ctypes.c_char_p
    buf = (ctypes.c_double * (3 * 16))()
    n = l.mcgvc_profile_collect_kinds(buf, 16)
    return [{"kernel": l.mcgvc_profile_kind_name(i).decode(), "ms": buf[3 * i], "flops": buf[3 * i + 1],
             "launches": int(buf[3 * i + 2])} for i in range(n)]
