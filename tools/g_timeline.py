"""Warm, in-order kernel list of ONE Generator forward (no_grad) and ONE forward+backward at batch B
(torch.profiler / CUPTI activity records as the clock): every launch in execution order with its stream,
start offset and duration -- the picture that says which launches sit on the critical path.

    python tools/g_timeline.py [--precision c8] [--batch 64] [--frames 64] [--out gpurun_out/g_timeline.md]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcgvc_loader  # noqa: E402


def capture(fn):
    from torch.profiler import ProfilerActivity, profile
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = []
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name.replace("(anonymous namespace)::", "").split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        evs.append((ev.time_range.start, ev.time_range.end, getattr(ev, "device_resource_id", 0), name))
    evs.sort()
    return evs


def table(title, evs):
    t0 = evs[0][0]
    span = max(e[1] for e in evs) - t0
    lines = ["## " + title, "", "span %.1f us, %d launches, summed %.1f us" % (span, len(evs), sum(e[1] - e[0] for e in evs)), "",
             "| # | start us | dur us | gap before us | stream | kernel |", "|---|---|---|---|---|---|"]
    last_end = {}
    for i, (s, e, st, name) in enumerate(evs):
        gap = s - last_end.get(st, s)
        last_end[st] = e
        lines.append("| %d | %.1f | %.1f | %.1f | %s | `%s` |" % (i, s - t0, e - s, gap, st, name[:80]))
    return "\n".join(lines) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="c8w")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "g_timeline.md"))
    args = ap.parse_args()
    pkg = mcgvc_loader.load()
    eng = pkg.engine
    eng.set_precision({"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "c8w": eng.PRECISION_C8W, "c8h": eng.PRECISION_C8H,
                       "mixed": eng.PRECISION_MIXED, "fast": eng.PRECISION_FAST}[args.precision])
    torch.manual_seed(0)
    G = pkg.Generator().to("cuda")
    x = torch.randn(args.batch, 80, args.frames, device="cuda")
    m = torch.ones_like(x)

    def fwd():
        with torch.no_grad():
            G(x, m)

    def fwd_bwd():
        G.zero_grad(set_to_none=True)
        G(x, m).sum().backward()

    text = "# Generator kernel timeline, batch %d x %d frames, %s mode\n\n" % (args.batch, args.frames, args.precision)
    text += table("forward (no_grad)", capture(fwd)) + "\n" + table("forward + backward", capture(fwd_bwd))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text)
    print(text[:6000])


if __name__ == "__main__":
    main()
