"""Quick A/B timer: ms per Generator forward+backward and per full train step at batch 64.
    python tools/quick_step.py [parity|c8|mixed|fast] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcgvc_loader  # noqa: E402
from bench import synthetic_batch_host  # noqa: E402

pkg = mcgvc_loader.load()
eng = pkg.engine
from maskcyclegan_vc_b200 import trainstep as ts  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "c8w": eng.PRECISION_C8W, "c8h": eng.PRECISION_C8H, "mixed": eng.PRECISION_MIXED,
                   "fast": eng.PRECISION_FAST}[mode])
dev = torch.device("cuda", 0)
models = ts.build_models(pkg.Generator, pkg.Discriminator, dev, seed=0)
g_opt, d_opt = ts.build_optimizers(models)
batch = [t.to(dev) for t in synthetic_batch_host(64, 64, seed=1234)]


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


G0 = models[0]


def gfb():
    G0.zero_grad(set_to_none=True)
    G0(batch[0], batch[1]).sum().backward()


def gf():
    with torch.no_grad():
        G0(batch[0], batch[1])


print("%s: G fwd %.3f ms | G fwd+bwd %.3f ms | train step %.2f ms"
      % (mode, timeit(gf, 20), timeit(gfb, 10), timeit(lambda: ts.train_step(models, g_opt, d_opt, batch), reps)), flush=True)
