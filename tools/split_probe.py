"""Probe: does running two independent half-batch Generator forwards on two streams (the memory-bound layer
kernels of one overlapping the tensor-core kernels of the other) beat one full-batch forward?
    python tools/split_probe.py [c8|c8h|parity]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcgvc_loader  # noqa: E402

pkg = mcgvc_loader.load()
eng = pkg.engine
mode = sys.argv[1] if len(sys.argv) > 1 else "c8"
eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "c8w": eng.PRECISION_C8W, "c8h": eng.PRECISION_C8H}[mode])
torch.manual_seed(0)
G = pkg.Generator().cuda()
D = pkg.Discriminator().cuda()
B = 64
x = torch.randn(B, 80, 64, device="cuda")
m = torch.ones_like(x)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def full(mod, *a):
    with torch.no_grad():
        return mod(*a)


def split(mod, *a):
    with torch.no_grad():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        h = B // 2
        with torch.cuda.stream(s1):
            y1 = mod(*[t[:h] for t in a])
        with torch.cuda.stream(s2):
            y2 = mod(*[t[h:] for t in a])
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        return y1, y2


print("%s  G fwd: full %.3f ms | two half-batch streams %.3f ms | two halves, one stream %.3f ms" % (
    mode, timeit(lambda: full(G, x, m)), timeit(lambda: split(G, x, m)),
    timeit(lambda: (full(G, x[:32], m[:32]), full(G, x[32:], m[32:])))))
print("%s  D fwd: full %.3f ms | two half-batch streams %.3f ms" % (mode, timeit(lambda: full(D, x)), timeit(lambda: split(D, x))))
