"""Kernel-level checks of the two tensor-core kernels through the C ABI (needs a B200).

For each case the same bf16 hi/lo operands go through (a) the tcgen05/TMA kernel, (b) the SIMT
checking kernel and (c) an fp64 torch evaluation of the tap sum; all three must agree.  Run as
`python tests/kernel_check.py` for a verbose report, or through pytest (tests/test_gpu_kernels.py).
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "maskcyclegan-vc_b200", "libmcgvc.so")


def load_lib():
    lib = ctypes.CDLL(LIB)
    lib.mcgvc_last_error.restype = ctypes.c_char_p
    return lib


def split_bf16(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def taps_array(taps):
    flat = []
    for t in taps:
        flat += list(t)
    return (ctypes.c_int8 * len(flat))(*flat)


def ref_conv(A, W, taps, oB, oY, oX):
    """A: [B,P,Y,X,C] float64, W: [T,N,C] float64 -> [oB,oY,oX,N]."""
    B, P, Y, X, C = A.shape
    pad = 16
    Ap = torch.zeros(B, P, Y + 2 * pad + oY, X + 2 * pad + oX, C, dtype=A.dtype, device=A.device)
    Ap[:, :, pad:pad + Y, pad:pad + X] = A
    out = torch.zeros(oB, oY, oX, W.shape[1], dtype=A.dtype, device=A.device)
    for (dx, dy, plane, w) in taps:
        win = Ap[:oB, plane, pad + dy:pad + dy + oY, pad + dx:pad + dx + oX]
        out += torch.einsum("byxc,nc->byxn", win, W[w])
    return out


def run_conv(lib, Ah, Al, Wh, Wl, taps, oB, oY, oX, nPass, backend, blockN, bias=None,
             addsrc=None, nSplit=None, out_shape=None, strides=None, ksplit=1):
    B, P, Y, X, C = Ah.shape
    T, N, K = Wh.shape
    lib.mcgvc_debug_set_conv_ksplit(ksplit)
    if nSplit is None:
        nSplit = N
        # split-K slices are ADDED into a zero-filled output (SIMT checker: plain store)
        fill = 0.0 if (ksplit > 1 and backend != 1) else float("nan")
        out = torch.full((oB, oY, oX, N), fill, device="cuda")
        sB, sY, sX, sNhi = oY * oX * N, oX * N, N, 0
    else:
        out = torch.full(out_shape, float("nan"), device="cuda")
        sB, sY, sX, sNhi = strides
    rc = lib.mcgvc_debug_conv(ptr(Ah), ptr(Al), C, X, Y, P, B, ptr(Wh), ptr(Wl), K, N, T, oX, oY,
                              oB, len(taps), taps_array(taps), ptr(out),
                              ctypes.c_longlong(sB), ctypes.c_longlong(sY), ctypes.c_longlong(sX),
                              nSplit, ctypes.c_longlong(sNhi), ptr(bias), ptr(addsrc), nPass,
                              backend, blockN, ctypes.c_void_p(0))
    lib.mcgvc_debug_set_conv_ksplit(1)
    if rc != 0:
        raise RuntimeError("mcgvc_debug_conv: " + lib.mcgvc_last_error().decode())
    torch.cuda.synchronize()
    return out


def relerr(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def conv_case(lib, name, B, P, Y, X, C, N, taps, oB, oY, oX, nPass, blockN, with_bias=True,
              with_add=False, seed=0, ksplit=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(B, P, Y, X, C, device="cuda", generator=g)
    T = max(t[3] for t in taps) + 1
    W = torch.randn(T, N, C, device="cuda", generator=g) * 0.05
    Ah, Al = split_bf16(A)
    Wh, Wl = split_bf16(W)
    bias = torch.randn(N, device="cuda", generator=g) if with_bias else None
    addsrc = torch.randn(oB, oY, oX, N, device="cuda", generator=g) if with_add else None
    if nPass == 3:
        Ae = Ah.double() + Al.double()
        We = Wh.double() + Wl.double()
        ref = ref_conv(Ae, We, taps, oB, oY, oX) - ref_conv(Al.double(), Wl.double(), taps, oB, oY, oX)
    else:
        ref = ref_conv(Ah.double(), Wh.double(), taps, oB, oY, oX)
    if bias is not None:
        ref = ref + bias.double()
    if addsrc is not None:
        ref = ref + addsrc.double()
    res = {}
    backends = [(1, "simt"), (0, "tc")]
    if N % 128 == 0:
        backends.append((2, "tc2"))   # CTA-pair kernel (cta_group::2)
    for backend, bname in backends:
        out = run_conv(lib, Ah, Al, Wh, Wl, taps, oB, oY, oX, nPass, backend, blockN, bias, addsrc, ksplit=ksplit)
        nan = torch.isnan(out).sum().item()
        res[bname] = (relerr(torch.nan_to_num(out), ref), nan)
    ok = all(e < 2e-5 and n == 0 for e, n in res.values())
    tc2 = f"| tc2 err={res['tc2'][0]:.2e} nan={res['tc2'][1]} " if "tc2" in res else ""
    if ksplit > 1:
        name = f"{name} kSplit={ksplit}"
    print(f"[conv ] {name:34s} nPass={nPass} blockN={blockN:3d} simt err={res['simt'][0]:.2e} nan={res['simt'][1]} "
          f"| tc err={res['tc'][0]:.2e} nan={res['tc'][1]} {tc2}{'OK' if ok else 'FAIL'}", flush=True)
    return ok


def split_c8(x, u, main_bf16=False):
    """C8 planes of a tensor: 16-bit hi (fp16 / bf16), e4m3(hi * 2^u), e4m3((x - hi) * 2^(u + 11 | 8))."""
    hi = x.to(torch.bfloat16 if main_bf16 else torch.float16)
    lo = x - hi.float()
    shift = 8 if main_bf16 else 11
    h8 = (hi.float() * 2.0 ** u).clamp(-448, 448).to(torch.float8_e4m3fn)
    l8 = (lo * 2.0 ** (u + shift)).clamp(-448, 448).to(torch.float8_e4m3fn)
    return hi.contiguous(), h8.contiguous(), l8.contiguous()


def conv_c8_case(lib, name, B, P, Y, X, C, N, taps, oB, oY, oX, blockN, main_bf16=False, with_bias=True,
                 with_add=False, seed=0, uA=1, uW=6, verbose=True):
    """16-bit main pass + two e4m3 correction passes (conv_c8.cu): tcgen05 pair kernel vs SIMT checker vs
    an fp64 evaluation of the same planes; also reports how far the scheme is from the exact product."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(B, P, Y, X, C, device="cuda", generator=g)
    T = max(t[3] for t in taps) + 1
    W = torch.randn(T, N, C, device="cuda", generator=g) * 0.05
    A16, A8h, A8l = split_c8(A, uA, main_bf16)
    W16, W8h, W8l = split_c8(W, uW, main_bf16)
    shift = 8 if main_bf16 else 11
    c2 = 2.0 ** -(uA + uW + shift)
    bias = torch.randn(N, device="cuda", generator=g) if with_bias else None
    addsrc = torch.randn(oB, oY, oX, N, device="cuda", generator=g) if with_add else None
    d = torch.float64
    ref = ref_conv(A16.to(d), W16.to(d), taps, oB, oY, oX) + c2 * (
        ref_conv(A8h.to(d), W8l.to(d), taps, oB, oY, oX) + ref_conv(A8l.to(d), W8h.to(d), taps, oB, oY, oX))
    exact = ref_conv(A.to(d), W.to(d), taps, oB, oY, oX)
    scheme_err = relerr(ref, exact)
    if bias is not None:
        ref = ref + bias.double()
    if addsrc is not None:
        ref = ref + addsrc.double()
    res = {}
    for backend, bname in ((1, "simt"), (2, "tc2")):
        out = torch.full((oB, oY, oX, N), float("nan"), device="cuda")
        rc = lib.mcgvc_debug_conv_c8(ptr(A16), ptr(A8h), ptr(A8l), C, X, Y, P, B, ptr(W16), ptr(W8h), ptr(W8l),
                                     C, N, T, oX, oY, oB, len(taps), taps_array(taps), ptr(out), ptr(bias),
                                     ptr(addsrc), 1 if main_bf16 else 0, ctypes.c_float(1.0), ctypes.c_float(c2),
                                     backend, blockN, ctypes.c_void_p(0))
        if rc != 0:
            raise RuntimeError("mcgvc_debug_conv_c8: " + lib.mcgvc_last_error().decode())
        torch.cuda.synchronize()
        res[bname] = (relerr(torch.nan_to_num(out), ref), torch.isnan(out).sum().item())
    ok = all(e < 2e-5 and n == 0 for e, n in res.values())
    if verbose:
        print(f"[c8   ] {name:34s} main={'bf16' if main_bf16 else 'fp16'} blockN={blockN:3d} simt err={res['simt'][0]:.2e} "
              f"nan={res['simt'][1]} | tc2 err={res['tc2'][0]:.2e} nan={res['tc2'][1]} | scheme vs exact product "
              f"{scheme_err:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def wgrad_c8_case(lib, name, zshape, xshape, taps, ztaps, pB, pY, pX, cTile, splitK, main_bf16=False, seed=0,
                  uZ=4, uX=1, zscale=0.1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    zB, zY, zX, N = zshape
    xB, xP, xY, xX, C = xshape
    Z = torch.randn(zB, 1, zY, zX, N, device="cuda", generator=g) * zscale
    Xa = torch.randn(xB, xP, xY, xX, C, device="cuda", generator=g)
    Z16, Z8h, Z8l = split_c8(Z, uZ, main_bf16)
    X16, X8h, X8l = split_c8(Xa, uX, main_bf16)
    shift = 8 if main_bf16 else 11
    c2 = 2.0 ** -(uZ + uX + shift)
    d = torch.float64
    ref = ref_wgrad(Z16.to(d), X16.to(d), taps, ztaps, pB, pY, pX) + c2 * (
        ref_wgrad(Z8h.to(d), X8l.to(d), taps, ztaps, pB, pY, pX) + ref_wgrad(Z8l.to(d), X8h.to(d), taps, ztaps, pB, pY, pX))
    exact = ref_wgrad(Z.to(d), Xa.to(d), taps, ztaps, pB, pY, pX)
    scheme_err = relerr(ref, exact)
    T = ref.shape[0]
    res = {}
    for backend, bname in ((1, "simt"), (2, "tc2")):
        dw = torch.zeros(T, N, C, device="cuda")
        rc = lib.mcgvc_debug_wgrad_c8(ptr(Z16), ptr(Z8h), ptr(Z8l), N, zX, zY, zB, ptr(X16), ptr(X8h), ptr(X8l), C, xX, xY,
                                      xP, xB, pX, pY, pB, len(taps), taps_array(taps), taps_array(ztaps), ptr(dw), cTile,
                                      splitK, 1 if main_bf16 else 0, ctypes.c_float(1.0), ctypes.c_float(c2), backend,
                                      ctypes.c_void_p(0))
        if rc != 0:
            raise RuntimeError("mcgvc_debug_wgrad_c8: " + lib.mcgvc_last_error().decode())
        torch.cuda.synchronize()
        res[bname] = relerr(dw, ref)
    ok = all(e < 2e-5 for e in res.values())
    print(f"[wg c8] {name:34s} main={'bf16' if main_bf16 else 'fp16'} cTile={cTile:3d} splitK={splitK} simt err={res['simt']:.2e} "
          f"| tc2 err={res['tc2']:.2e} | scheme vs exact product {scheme_err:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def wgrad_c8_cases(lib):
    ok = True
    z0 = [(0, 0, 0, 0)]
    ok &= wgrad_c8_case(lib, "1tap B2 Y4 X16 N256 C256", (2, 4, 16, 256), (2, 1, 4, 16, 256), z0, z0, 2, 4, 16, 256, 1)
    ok &= wgrad_c8_case(lib, "5x5 s1 N512 C256", (2, 20, 16, 512), (2, 1, 20, 16, 256), taps_5x5_s1(), z0 * 25, 2, 20, 16, 256, 2)
    ok &= wgrad_c8_case(lib, "5x5 s1 N512 C256 bf16 main", (2, 20, 16, 512), (2, 1, 20, 16, 256), taps_5x5_s1(), z0 * 25, 2, 20, 16, 256, 2, main_bf16=True)
    ok &= wgrad_c8_case(lib, "5x5 s2 parity N256 C256 odd", (3, 10, 9, 256), (3, 4, 10, 9, 256), taps_kxk_s2(5, 2), z0 * 25, 3, 10, 9, 256, 1)
    ok &= wgrad_c8_case(lib, "ztaps (1dto2d style) 20 taps", (4, 20, 16, 256), (4, 1, 1, 16, 256),
                        [(0, 0, 0, h) for h in range(20)], [(0, h, 0, 0) for h in range(20)], 4, 1, 16, 256, 1)
    ok &= wgrad_c8_case(lib, "B16 Y40 X32 N512 C256 splitK4", (16, 40, 32, 512), (16, 1, 40, 32, 256), taps_5x5_s1(), z0 * 25, 16, 40, 32, 256, 4)
    # channel tile 128: each CTA's half is 64 channels -> 64-byte rows, 64B swizzle, MN-major
    ok &= wgrad_c8_case(lib, "1tap N256 C128 (cTile 128)", (2, 4, 16, 256), (2, 1, 4, 16, 128), z0, z0, 2, 4, 16, 128, 1)
    ok &= wgrad_c8_case(lib, "5x5 s2 parity N512 C128 (cTile 128)", (2, 20, 16, 512), (2, 4, 20, 16, 128), taps_kxk_s2(5, 2), z0 * 25, 2, 20, 16, 128, 2)
    return ok


def c8_cases(lib):
    ok = True
    for bn in (128, 256):
        ok &= conv_c8_case(lib, "1tap B2 Y4 X16 C64 N256", 2, 1, 4, 16, 64, 256, [(0, 0, 0, 0)], 2, 4, 16, bn)
        ok &= conv_c8_case(lib, "5x5 s1 B2 Y20 X16 C128 N256", 2, 1, 20, 16, 128, 256, taps_5x5_s1(), 2, 20, 16, bn, with_add=True)
    ok &= conv_c8_case(lib, "5x5 s1 bf16 main", 2, 1, 20, 16, 128, 256, taps_5x5_s1(), 2, 20, 16, 256, main_bf16=True)
    ok &= conv_c8_case(lib, "5x5 s1 odd B3 Y5 X17 C64 N128", 3, 1, 5, 17, 64, 128, taps_5x5_s1(), 3, 5, 17, 128)
    ok &= conv_c8_case(lib, "5x5 s2 parity B2 40x32->20x16", 2, 4, 20, 16, 128, 256, taps_kxk_s2(5, 2), 2, 20, 16, 256)
    ok &= conv_c8_case(lib, "5x5 s1 B16 Y40 X32 C128 N512 (waves)", 16, 1, 40, 32, 128, 512, taps_5x5_s1(), 16, 40, 32, 256)
    ok &= conv_c8_case(lib, "same, blockN 128 (2 TMEM buffers)", 16, 1, 40, 32, 128, 512, taps_5x5_s1(), 16, 40, 32, 128)
    return ok


def taps_5x5_s1():
    return [(kw - 2, kh - 2, 0, kh * 5 + kw) for kh in range(5) for kw in range(5)]


def taps_kxk_s2(k, pad):
    taps = []
    for kh in range(k):
        for kw in range(k):
            oy, ox = kh - pad, kw - pad
            ph, pw = oy % 2, ox % 2
            taps.append(((ox - pw) // 2, (oy - ph) // 2, ph * 2 + pw, kh * k + kw))
    return taps


def ref_wgrad(Z, Xa, taps, ztaps, pB, pY, pX):
    """Z: [B,1,Y,X,N], Xa: [B,P,Y,X,C] -> [T,N,C]."""
    pad = 16

    def padded(A):
        B, P, Y, X, C = A.shape
        Ap = torch.zeros(max(B, pB), P, Y + 2 * pad + pY, X + 2 * pad + pX, C, dtype=A.dtype, device=A.device)
        Ap[:B, :, pad:pad + Y, pad:pad + X] = A
        return Ap

    Zp, Xp = padded(Z), padded(Xa)
    T = max(t[3] for t in taps) + 1
    out = torch.zeros(T, Z.shape[-1], Xa.shape[-1], dtype=Z.dtype, device=Z.device)
    for (dx, dy, plane, w), (zdx, zdy, _, _) in zip(taps, ztaps):
        zw = Zp[:pB, 0, pad + zdy:pad + zdy + pY, pad + zdx:pad + zdx + pX]
        xw = Xp[:pB, plane, pad + dy:pad + dy + pY, pad + dx:pad + dx + pX]
        out[w] += torch.einsum("byxn,byxc->nc", zw, xw)
    return out


def wgrad_case(lib, name, zshape, xshape, taps, ztaps, pB, pY, pX, nPass, cTile, splitK, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    zB, zY, zX, N = zshape
    xB, xP, xY, xX, C = xshape
    Z = torch.randn(zB, 1, zY, zX, N, device="cuda", generator=g) * 0.1
    Xa = torch.randn(xB, xP, xY, xX, C, device="cuda", generator=g)
    Zh, Zl = split_bf16(Z)
    Xh, Xl = split_bf16(Xa)
    if nPass == 3:
        ref = ref_wgrad(Zh.double() + Zl.double(), Xh.double() + Xl.double(), taps, ztaps, pB, pY, pX) - \
            ref_wgrad(Zl.double(), Xl.double(), taps, ztaps, pB, pY, pX)
    else:
        ref = ref_wgrad(Zh.double(), Xh.double(), taps, ztaps, pB, pY, pX)
    T = ref.shape[0]
    res = {}
    backends = [(1, "simt"), (0, "tc")]
    if N % 256 == 0 and cTile >= 128:
        backends.append((2, "tc2"))   # CTA-pair kernel
    for backend, bname in backends:
        dw = torch.zeros(T, N, C, device="cuda")
        rc = lib.mcgvc_debug_wgrad(ptr(Zh), ptr(Zl), N, zX, zY, zB, ptr(Xh), ptr(Xl), C, xX, xY, xP,
                                   xB, pX, pY, pB, len(taps), taps_array(taps), taps_array(ztaps),
                                   ptr(dw), cTile, splitK, nPass, backend, ctypes.c_void_p(0))
        if rc != 0:
            raise RuntimeError("mcgvc_debug_wgrad: " + lib.mcgvc_last_error().decode())
        torch.cuda.synchronize()
        res[bname] = relerr(dw, ref)
    ok = all(e < 2e-5 for e in res.values())
    tc2 = f"| tc2 err={res['tc2']:.2e} " if "tc2" in res else ""
    print(f"[wgrad] {name:34s} nPass={nPass} cTile={cTile:3d} splitK={splitK} simt err={res['simt']:.2e} "
          f"| tc err={res['tc']:.2e} {tc2}{'OK' if ok else 'FAIL'}", flush=True)
    return ok


def all_cases(lib):
    ok = True
    z0 = [(0, 0, 0, 0)]
    # --- conv: 1 tap, plain GEMM shapes
    for nPass in (1, 3):
        for bn in (64, 128, 256):
            ok &= conv_case(lib, "1tap B2 Y4 X16 C64 N256", 2, 1, 4, 16, 64, 256, [(0, 0, 0, 0)], 2, 4, 16, nPass, bn)
    ok &= conv_case(lib, "1tap C256 (4 kblocks)", 2, 1, 4, 16, 256, 128, [(0, 0, 0, 0)], 2, 4, 16, 3, 128)
    ok &= conv_case(lib, "1tap many tiles B8 Y20 X16", 8, 1, 20, 16, 128, 256, [(0, 0, 0, 0)], 8, 20, 16, 3, 128)
    # --- conv: 5x5 stride 1 with padding (up blocks / dgrad)
    ok &= conv_case(lib, "5x5 s1 B2 Y20 X16 C128 N256", 2, 1, 20, 16, 128, 256, taps_5x5_s1(), 2, 20, 16, 3, 128, with_add=True)
    ok &= conv_case(lib, "5x5 s1 odd B3 Y5 X17 C64", 3, 1, 5, 17, 64, 128, taps_5x5_s1(), 3, 5, 17, 3, 128)
    ok &= conv_case(lib, "5x5 s1 bf16 N256 bn256", 2, 1, 20, 16, 128, 256, taps_5x5_s1(), 2, 20, 16, 1, 256)
    # --- conv: stride 2 through parity planes
    ok &= conv_case(lib, "5x5 s2 parity B2 40x32->20x16", 2, 4, 20, 16, 128, 256, taps_kxk_s2(5, 2), 2, 20, 16, 3, 128)
    ok &= conv_case(lib, "3x3 s2 parity odd X", 2, 4, 10, 9, 64, 128, taps_kxk_s2(3, 1), 2, 10, 9, 3, 128)
    # --- conv: 20 taps down Y with oY=1 (2D->1D flatten)
    ok &= conv_case(lib, "2dto1d 20 taps oY=1", 4, 1, 20, 16, 256, 256, [(0, h, 0, h) for h in range(20)], 4, 1, 16, 3, 128)
    # --- conv: 1D k=3
    ok &= conv_case(lib, "1D k3 B8 L16 C256 N1024", 8, 1, 1, 16, 256, 1024, [(-1, 0, 0, 0), (0, 0, 0, 1), (1, 0, 0, 2)], 8, 1, 16, 3, 128)
    # --- conv: split-K work items (K slices added into a zero-filled output with red.global.add)
    ok &= conv_case(lib, "5x5 s1 B2 Y20 X16 C128 N256", 2, 1, 20, 16, 128, 256, taps_5x5_s1(), 2, 20, 16, 3, 128, with_add=True, ksplit=3)
    ok &= conv_case(lib, "5x5 s1 B2 Y20 X16 C128 N256", 2, 1, 20, 16, 128, 256, taps_5x5_s1(), 2, 20, 16, 3, 256, ksplit=7)
    ok &= conv_case(lib, "5x5 s1 bf16 odd B3 Y5 X17 C64", 3, 1, 5, 17, 64, 128, taps_5x5_s1(), 3, 5, 17, 1, 128, ksplit=5)
    ok &= conv_case(lib, "3x3 s2 parity C256 (4 kblocks/tap)", 2, 4, 10, 9, 256, 128, taps_kxk_s2(3, 1), 2, 10, 9, 3, 64, ksplit=8)
    ok &= conv_case(lib, "1D k3 B8 L16 C256 N1024", 8, 1, 1, 16, 256, 1024, [(-1, 0, 0, 0), (0, 0, 0, 1), (1, 0, 0, 2)], 8, 1, 16, 3, 128, ksplit=2)
    # --- wgrad
    for nPass in (1, 3):
        for ct in (64, 128, 256):
            ok &= wgrad_case(lib, "1tap B2 Y4 X16 N128 C256", (2, 4, 16, 128), (2, 1, 4, 16, 256), z0, z0, 2, 4, 16, nPass, ct, 1)
    ok &= wgrad_case(lib, "1tap splitK3 B8 Y20 X16", (8, 20, 16, 256), (8, 1, 20, 16, 128), z0, z0, 8, 20, 16, 3, 128, 3)
    ok &= wgrad_case(lib, "5x5 s1 B2 Y20 X16", (2, 20, 16, 128), (2, 1, 20, 16, 128), taps_5x5_s1(), z0 * 25, 2, 20, 16, 3, 128, 2)
    ok &= wgrad_case(lib, "5x5 s2 parity", (2, 20, 16, 128), (2, 4, 20, 16, 64), taps_kxk_s2(5, 2), z0 * 25, 2, 20, 16, 3, 64, 1)
    ok &= wgrad_case(lib, "odd B3 Y5 X17 3x3 s2", (3, 5, 9, 128), (3, 4, 5, 9, 64), taps_kxk_s2(3, 1), z0 * 9, 3, 5, 9, 3, 64, 2)
    for nPass in (1, 3):
        for ct in (128, 256):
            ok &= wgrad_case(lib, "pair 5x5 s1 N512 C256", (2, 20, 16, 512), (2, 1, 20, 16, 256), taps_5x5_s1(), z0 * 25, 2, 20, 16, nPass, ct, 2)
    ok &= wgrad_case(lib, "pair 5x5 s2 parity N256 C128 odd", (3, 10, 9, 256), (3, 4, 10, 9, 128), taps_kxk_s2(5, 2), z0 * 25, 3, 10, 9, 3, 128, 1)
    ok &= wgrad_case(lib, "ztaps (1dto2d style) 20 taps", (4, 20, 16, 256), (4, 1, 1, 16, 256),
                     [(0, 0, 0, h) for h in range(20)], [(0, h, 0, 0) for h in range(20)], 4, 1, 16, 3, 128, 1)
    return ok


if __name__ == "__main__":
    lib = load_lib()
    if len(sys.argv) > 1 and sys.argv[1] == "c8":
        ok = c8_cases(lib)
        ok &= wgrad_c8_cases(lib)
        print("KERNEL_CHECK_C8", "PASS" if ok else "FAIL")
        sys.exit(0 if ok else 1)
    ok = all_cases(lib)
    print("KERNEL_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
