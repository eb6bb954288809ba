"""Warm, in-context kernel timeline of one full train step (no ncu: kernels run back to back at the
clocks and cache state of the real step, with the weight-gradient side stream overlapping).

    python tools/timeline.py [--batch 64] [--precision parity] [--out gpurun_out/timeline.md]

Uses torch.profiler (kineto / CUPTI activity records, bundled with torch) purely as a clock: per
kernel name it prints launches, summed device time and share, plus the step's wall time on the GPU
and the busy time of each stream.  ncu's launch list (profiles/*_launches.csv.gz) serialises the
kernels and flushes caches per launch, which over-states every small kernel; this is the list that
says where the step's time really goes.
"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcgvc_loader  # noqa: E402
from bench import synthetic_batch_host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="parity", choices=["parity", "c8", "mixed", "fast"])
    ap.add_argument("--lean", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.md"))
    args = ap.parse_args()
    pkg = mcgvc_loader.load()
    eng = pkg.engine
    from maskcyclegan_vc_b200 import trainstep as ts
    eng.lib()
    eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "mixed": eng.PRECISION_MIXED, "fast": eng.PRECISION_FAST}[args.precision])
    pkg.set_lean(bool(args.lean))
    dev = torch.device("cuda", 0)
    models = ts.build_models(pkg.Generator, pkg.Discriminator, dev, seed=0)
    g_opt, d_opt = ts.build_optimizers(models)
    batch = [t.to(dev) for t in synthetic_batch_host(args.batch, 64, seed=1234)]
    for _ in range(3):
        ts.train_step(models, g_opt, d_opt, batch)
    torch.cuda.synchronize()

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            ts.train_step(models, g_opt, d_opt, batch)
        torch.cuda.synchronize()

    per = collections.defaultdict(lambda: [0, 0.0])
    streams = collections.defaultdict(float)
    t_min, t_max = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        dur = float(ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total)
        name = ev.name.replace("(anonymous namespace)::", "")
        short = name.split("(")[0]
        if short.startswith("void "):
            short = short[5:]
        per[short][0] += 1
        per[short][1] += dur
        tr = ev.time_range
        t_min = tr.start if t_min is None else min(t_min, tr.start)
        t_max = tr.end if t_max is None else max(t_max, tr.end)
        streams[getattr(ev, "device_resource_id", 0)] += dur
    total = sum(v[1] for v in per.values())
    span = (t_max - t_min) if t_min is not None else 0.0
    lines = ["# Warm kernel timeline: %d train steps at batch %d, %s mode, lean=%d" % (args.steps, args.batch, args.precision, args.lean), "",
             "GPU span %.2f ms per step; summed kernel time %.2f ms per step (sum > span = stream overlap, sum < span = gaps)" %
             (span / 1e3 / args.steps, total / 1e3 / args.steps),
             "busy ms per step by stream: " + ", ".join("%s: %.2f" % (k, v / 1e3 / args.steps) for k, v in sorted(streams.items(), key=lambda kv: -kv[1])), "",
             "| kernel | launches/step | ms/step | share of summed | avg us |", "|---|---|---|---|---|"]
    for name, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %.1f | %.3f | %.1f%% | %.1f |" % (name[:90], n / args.steps, us / 1e3 / args.steps, 100.0 * us / total, us / n))
    text = "\n".join(lines) + "\n"
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text)
    print(text)


if __name__ == "__main__":
    main()
