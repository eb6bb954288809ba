"""Measures BASELINE.json configs 1, 2, 3 and 5 (config 4 is bench.py's headline) and prints one
JSON object: engine on cuda:0 (default C8W mode, or argv[1]) next to the CPU oracle port where that is cheap."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import maskcyclegan_oracle as O  # noqa: E402  (CPU baseline leg only)
import mcgvc_loader  # noqa: E402

pkg = mcgvc_loader.load()
from maskcyclegan_vc_b200 import trainstep as ts  # noqa: E402

eng = pkg.engine


def gpu_time(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_time(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "c8w"
    eng.lib()
    eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "c8w": eng.PRECISION_C8W, "c8h": eng.PRECISION_C8H,
                       "mixed": eng.PRECISION_MIXED, "fast": eng.PRECISION_FAST}[mode])
    out = {"cores": os.cpu_count(), "precision": mode}
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    G, D = pkg.Generator().cuda(), pkg.Discriminator().cuda()
    torch.manual_seed(0)
    gs, ds = O.build_generator_state(), O.build_discriminator_state()

    # config 1: G A2B forward, B=1, 80x64, no_grad
    x, m, _, _ = O.synthetic_batch(1, 64, seed=1234)
    xc, mc = x.cuda(), m.cuda()
    with torch.no_grad():
        cpu_ms = cpu_time(lambda: O.generator_forward(gs, x, m), reps=10)
        gpu_ms = gpu_time(lambda: G(xc, mc), reps=20)
    out["config1_G_fwd_B1"] = {"cpu_ms": cpu_ms, "cpu_frames_s": 64 / cpu_ms * 1e3, "gpu_ms": gpu_ms, "gpu_frames_s": 64 / gpu_ms * 1e3}

    # config 2: G + D adversarial fwd+bwd, B=1
    gso = {k: v.clone().requires_grad_(True) for k, v in gs.items()}
    dso = {k: v.clone().requires_grad_(True) for k, v in ds.items()}

    def cpu_adv():
        for v in list(gso.values()) + list(dso.values()):
            v.grad = None
        torch.mean((1 - O.discriminator_forward(dso, O.generator_forward(gso, x, m))) ** 2).backward()

    def gpu_adv():
        G.zero_grad(set_to_none=True)
        D.zero_grad(set_to_none=True)
        torch.mean((1 - D(G(xc, mc))) ** 2).backward()

    cpu_ms = cpu_time(cpu_adv, reps=5)
    gpu_ms = gpu_time(gpu_adv, reps=20)
    out["config2_adv_fwd_bwd_B1"] = {"cpu_ms": cpu_ms, "cpu_frames_s": 64 / cpu_ms * 1e3, "gpu_ms": gpu_ms, "gpu_frames_s": 64 / gpu_ms * 1e3}

    # config 3: full train step, B=16
    models = ts.build_models(pkg.Generator, pkg.Discriminator, torch.device("cuda"), seed=0)
    g_opt, d_opt = ts.build_optimizers(models)
    batch = [t.cuda() for t in O.synthetic_batch(16, 64, seed=1234)]
    gpu_ms = gpu_time(lambda: ts.train_step(models, g_opt, d_opt, batch), warm=3, reps=5)
    out["config3_full_step_B16"] = {"gpu_ms": gpu_ms, "gpu_frames_s": 16 * 64 / gpu_ms * 1e3}
    del models, g_opt, d_opt, batch
    torch.cuda.empty_cache()

    # config 5: G inference, B=8, 80x512
    x5, m5, _, _ = O.synthetic_batch(8, 512, seed=7)
    x5c, m5c = x5.cuda(), m5.cuda()
    with torch.no_grad():
        gpu_ms = gpu_time(lambda: G(x5c, m5c), reps=10)
        x51, m51 = x5[:1], m5[:1]
        cpu_ms = cpu_time(lambda: O.generator_forward(gs, x51, m51), reps=3)
        err = ((G(x5c, m5c).cpu() - O.generator_forward(gs, x5, m5)).norm() / O.generator_forward(gs, x5, m5).norm()).item()
    out["config5_G_infer_B8_T512"] = {"gpu_ms": gpu_ms, "gpu_frames_s": 8 * 512 / gpu_ms * 1e3, "cpu_ms_B1": cpu_ms,
                                      "cpu_frames_s_B1": 512 / cpu_ms * 1e3, "rel_err_vs_oracle": err,
                                      "gpu_tflops_algorithmic": 8 * 8 * 19.676266496e9 / (gpu_ms * 1e-3) / 1e12}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
