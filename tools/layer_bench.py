"""Kernel-level timing of the tensor-core kernels on the big layer shapes (CUDA events, warm, L2
flushed between reps by the operand sizes themselves: > 126 MB touched per launch for batch 64).
    python tools/layer_bench.py [B]
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_check as kc  # noqa: E402

lib = kc.load_lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def time_fn(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_layer(name, P, Y, X, C, N, taps, oY, oX):
    A = torch.randn(B, P, Y, X, C, device="cuda")
    T = max(t[3] for t in taps) + 1
    W = torch.randn(T, N, C, device="cuda") * 0.05
    Ah, Al = kc.split_bf16(A)
    Wh, Wl = kc.split_bf16(W)
    del A, W
    flops = 2.0 * B * oY * oX * N * len(taps) * C
    # 16-bit main pass + two e4m3 correction passes (conv_c8.cu), valid numerics
    A = Ah.float() + Al.float()
    W = Wh.float() + Wl.float()
    A16, A8h, A8l = kc.split_c8(A, 1)
    W16, W8h, W8l = kc.split_c8(W, 6)
    del A, W
    out = torch.empty(B, oY, oX, N, device="cuda")
    row = []
    for bn in (128, 256):
        if N % bn:
            continue

        def run_c8():
            rc = lib.mcgvc_debug_conv_c8(kc.ptr(A16), kc.ptr(A8h), kc.ptr(A8l), C, X, Y, P, B, kc.ptr(W16), kc.ptr(W8h),
                                         kc.ptr(W8l), C, N, T, oX, oY, B, len(taps), kc.taps_array(taps), kc.ptr(out),
                                         kc.ptr(None), kc.ptr(None), 0, ctypes.c_float(1.0), ctypes.c_float(2.0 ** -18),
                                         2, bn, ctypes.c_void_p(0))
            assert rc == 0, lib.mcgvc_last_error()
        ms = time_fn(run_c8)
        row.append("pair/%d %.0fus %.0fTF" % (bn, ms * 1e3, flops / ms / 1e9))
    print("[conv ] %-10s C8 (f16 + 2x e4m3)  %s" % (name, " | ".join(row)), flush=True)
    del A16, A8h, A8l, W16, W8h, W8l, out
    for nPass in (3, 1):
        row = []
        for backend, bn, label in ((0, 128, "1cta/128"), (0, 256, "1cta/256"), (2, 128, "pair/128"), (2, 256, "pair/256")):
            if N % bn:
                continue
            ms = time_fn(lambda: kc.run_conv(lib, Ah, Al, Wh, Wl, taps, B, oY, oX, nPass, backend, bn))
            row.append("%s %.0fus %.0fTF" % (label, ms * 1e3, flops / ms / 1e9))
        print("[conv ] %-10s nPass=%d  %s" % (name, nPass, " | ".join(row)), flush=True)


def wgrad_layer(name, zY, zX, N, xP, xY, xX, C, taps):
    Z = torch.randn(B, 1, zY, zX, N, device="cuda") * 0.1
    Xa = torch.randn(B, xP, xY, xX, C, device="cuda")
    Zh, Zl = kc.split_bf16(Z)
    Xh, Xl = kc.split_bf16(Xa)
    del Z, Xa
    T = max(t[3] for t in taps) + 1
    dw = torch.zeros(T, N, C, device="cuda")
    zt = [(0, 0, 0, 0)] * len(taps)
    flops = 2.0 * B * zY * zX * N * C * len(taps)
    posTiles = B * zY * zX // 64

    def run(backend, ct, sk, nPass):
        rc = lib.mcgvc_debug_wgrad(kc.ptr(Zh), kc.ptr(Zl), N, zX, zY, B, kc.ptr(Xh), kc.ptr(Xl), C, xX, xY, xP, B,
                                   zX, zY, B, len(taps), kc.taps_array(taps), kc.taps_array(zt), kc.ptr(dw), ct, sk,
                                   nPass, backend, ctypes.c_void_p(0))
        assert rc == 0, lib.mcgvc_last_error()

    for nPass in (3, 1):
        row = []
        for backend, ct, label in ((0, 128, "1cta/128"), (2, 128, "pair/128"), (2, 256, "pair/256")):
            if C % ct or (backend == 2 and N % 256):
                continue
            units = len(taps) * (N // 128) * (C // ct)
            best = None
            for sk in sorted(set([max(1, (148 * k) // units) for k in (2, 3, 4, 6, 8)])):
                if sk > posTiles // 8:
                    continue
                ms = time_fn(lambda: run(backend, ct, sk, nPass), reps=3)
                if best is None or ms < best[0]:
                    best = (ms, sk)
            row.append("%s sk=%d %.0fus %.0fTF" % (label, best[1], best[0] * 1e3, flops / best[0] / 1e9))
        print("[wgrad] %-10s nPass=%d  %s" % (name, nPass, " | ".join(row)), flush=True)


if __name__ == "__main__":
    s1 = kc.taps_5x5_s1()
    s2 = kc.taps_kxk_s2(5, 2)
    conv_layer("up2", 1, 40, 32, 256, 512, s1, 40, 32)
    conv_layer("up1", 1, 20, 16, 256, 1024, s1, 20, 16)
    if len(sys.argv) > 2 and sys.argv[2] == "conv":
        sys.exit(0)
    conv_layer("ds1", 4, 40, 32, 128, 512, s2, 40, 32)
    conv_layer("ds2", 4, 20, 16, 256, 512, s2, 20, 16)
    conv_layer("D.ds2", 4, 20, 16, 256, 512, kc.taps_kxk_s2(3, 1), 20, 16)
    wgrad_layer("up2", 40, 32, 512, 1, 40, 32, 256, s1)
    wgrad_layer("up1", 20, 16, 1024, 1, 20, 16, 256, s1)
    wgrad_layer("ds1", 40, 32, 512, 4, 40, 32, 128, s2)
    wgrad_layer("ds2", 20, 16, 512, 4, 20, 16, 256, s2)
