"""Headline benchmark: mel-frames/sec of the full MaskCycleGAN-VC train step (train.py:186-299:
10 Generator + 12 Discriminator forwards, two backwards, two Adam steps, gradient all-reduce when
world > 1) on synthetic 80x64 mel batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
`--impl reference` times the reference's own modules (oracle/_ref, staged by oracle/build_ref.sh) on
the host CPU through the same train-step composition; the oracle port stands in only if oracle/_ref
did not travel with the snapshot.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mel-frames/sec full CycleGAN train step (2G+2D fwd+bwd) 80x64"
UNIT = "mel-frames/s"
T_FRAMES = 64


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
def synthetic_batch_host(B, T, seed, max_mask_len=25):
    """real_A/B ~ N(0,1); FIF masks as dataset/vc_dataset.py:51-55."""
    g = torch.Generator().manual_seed(seed)
    real_A = torch.randn(B, 80, T, generator=g)
    real_B = torch.randn(B, 80, T, generator=g)
    masks = []
    for _ in range(2):
        m = torch.ones(B, 80, T)
        for b in range(B):
            size = int(torch.randint(0, max_mask_len, (1,), generator=g))
            start = int(torch.randint(0, T - size, (1,), generator=g))
            m[b, :, start:start + size] = 0.0
        masks.append(m)
    return real_A, masks[0], real_B, masks[1]


class ClockSampler:
    """nvidia-smi sampling DURING the timed regions (B200_PROFILING.md clocks line)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.proc = None
        self.path = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.path = tempfile.NamedTemporaryFile(prefix="mcgvc_clk_", suffix=".csv", delete=False).name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", uuid, "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx, pw = [], set(), None, []
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) < 7:
                        continue
                    try:
                        sm.append(float(parts[0]))
                        mx = float(parts[1])
                        pw.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # under load = top half of samples (the sampler also sees the idle edges)
            top = sorted(sm)[len(sm) // 2:]
            out.update({"sm_mhz": statistics.median(top), "sm_max_mhz": mx, "reasons": sorted(reasons),
                        "samples": len(sm), "power_w_max": max(pw) if pw else None})
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm.  The reference's own Generator / Discriminator modules (oracle/_ref, staged by
# oracle/build_ref.sh; kind "reference") when they travelled with the snapshot, else the oracle port
# (kind "port").  The loop body is trainstep.train_step, the restatement of train.py:186-299 that
# also drives the engine (loaded straight from its file: the engine package is NOT imported here).
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _load_trainstep():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "mcgvc_trainstep_only", os.path.join(ROOT, "maskcyclegan-vc_b200", "trainstep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cpu_models():
    """(kind, models, step_fn(models, g_opt, d_opt, batch)) on the host CPU."""
    ts = _load_trainstep()
    if os.path.exists(os.path.join(REF_DIR, "mask_cyclegan_vc", "model.py")):
        sys.path.insert(0, REF_DIR)
        from mask_cyclegan_vc.model import Discriminator, Generator   # the UNMODIFIED reference modules
        kind = "reference"
        what = "reference mask_cyclegan_vc/model.py modules (oracle/_ref) + torch.optim.Adam"
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import maskcyclegan_oracle as O
        Generator, Discriminator = O.OracleGenerator, O.OracleDiscriminator
        kind = "port"
        what = "oracle port of model.py (oracle/_ref not staged) + torch.optim.Adam"
    models = ts.build_models(Generator, Discriminator, torch.device("cpu"), seed=0)
    g_opt, d_opt = ts.build_optimizers(models)
    return kind, what, models, g_opt, d_opt, ts.train_step


def _cpu_step_timer(B, seed=1234):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, what, models, g_opt, d_opt, step = _cpu_models()
    batch = list(synthetic_batch_host(B, T_FRAMES, seed=seed))

    def one():
        t0 = time.perf_counter()
        g, d = step(models, g_opt, d_opt, batch)
        float(g), float(d)                       # train.py:302-304 reads both losses every step
        return time.perf_counter() - t0
    return kind, what, cores, one


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the full train step on all host
    threads.  Same metric / config line as the engine arm; each step is a bounded SAMPLE of the
    batch-64 workload -- batch 16 (BASELINE configs[2]), or batch 4 if the first step shows that
    W + K steps at batch 16 would not finish within ~5 minutes on this box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = max(args.steps, 1), max(args.warmup, 1)
    B = args.ref_batch
    kind, what, cores, one = _cpu_step_timer(B)
    t_first = one()
    done_warm = 1
    if B > 4 and t_first * (K + W) > args.ref_budget_s:
        B = 4
        kind, what, cores, one = _cpu_step_timer(B)
        t_first = one()
    for _ in range(W - done_warm):
        one()
    times = [one() for _ in range(K)]
    dt = sum(times) / K
    value = B * T_FRAMES / dt
    sample = ("%s; fp32, %d host threads; each step = the full train step (train.py:186-299) at batch %d, a bounded "
              "sample of the batch-%d-per-GPU workload; %d warm-up + %d timed steps, %.2f s/step"
              % (what, cores, B, args.batch, W, K, dt))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args.batch, max(args.gpus, 1), args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(B, steps):
    """The same CPU arm, timed beside the engine on rank 0 at N = 1: 1 warm-up + `steps` timed steps."""
    kind, what, cores, one = _cpu_step_timer(B)
    one()
    dt = sum(one() for _ in range(steps)) / steps
    return {"value": B * T_FRAMES / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%s; full train step at batch %d (bounded sample of the batch-64 workload), fp32, %d host threads, "
                      "1 warm-up + %d timed steps, %.2f s/step" % (what, B, cores, steps, dt)}


def make_config(B, world, args):
    """`config` of the JSON line -- shared by the engine arm and the CPU reference arm."""
    return {"workload": "full MaskCycleGAN train step (train.py:186-299: 10 G fwd + 12 D fwd, 2 backward, 2 Adam), "
                        "batch %d per GPU, 80x%d mel (BASELINE configs[3])" % (B, T_FRAMES),
            "batch_per_gpu": B, "global_batch": B * world, "frames": T_FRAMES,
            "parallelism": "dp%d" % world,
            "l2": "inputs larger than L2: ~%.1f GB of activations touched per step vs 126 MB L2" % (24.0 * B / 64.0)}


# ------------------------------------------------------------------------------------------------
def timed_steps(step_fn, n, world):
    """barrier + synchronize on both sides, CUDA events on the current stream, max over ranks."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step_fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling)")
    ap.add_argument("--impl", default="engine")
    ap.add_argument("--precision", default="c8w", choices=["parity", "c8", "c8w", "c8h", "mixed", "fast"])
    ap.add_argument("--lean", type=int, default=0)
    ap.add_argument("--optimizer", default="torch", choices=["torch", "fused"],
                    help="torch = torch.optim.Adam as in train.py:119-122; fused = one-kernel Adam on the flat buffers")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--fast-steps", type=int, default=3, help="extra steps in the other precision mode (0 = skip)")
    ap.add_argument("--cpu-batch", type=int, default=16, help="batch of the in-line cpu_baseline sample (configs[2])")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=16, help="batch of one --impl reference step (configs[2])")
    ap.add_argument("--ref-budget-s", type=float, default=330.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import mcgvc_loader
    pkg = mcgvc_loader.load()
    eng = pkg.engine
    from maskcyclegan_vc_b200 import trainstep as ts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_out = sys.stdout
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner (and warnings) to fd 1; rank 0's stdout must carry ONE JSON line:
        # everything else written to fd 1 during the run goes to stderr, the line itself to the saved fd
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    B = args.batch
    eng.lib()
    eng.set_backend(eng.BACKEND_TCGEN05)
    modes = {"parity": eng.PRECISION_PARITY, "c8": eng.PRECISION_C8, "c8w": eng.PRECISION_C8W, "c8h": eng.PRECISION_C8H,
             "mixed": eng.PRECISION_MIXED, "fast": eng.PRECISION_FAST}
    mode_note = {"parity": "split-bf16 x3 in every GEMM (fwd, dgrad, wgrad): outputs and gradients within 1e-3 of the fp32 reference",
                 "c8": "fp16 main pass + two e4m3 correction passes in every GEMM but the stems/heads (those: split-bf16 x3), "
                       "2 MMA units per MAC: outputs and gradients within 1e-3 of the fp32 reference",
                 "c8w": "forward and data-gradient GEMMs exactly as c8; the weight-gradient GEMMs of the c8 layers (leaves of the backward "
                        "graph: their rounding is not propagated) run ONE fp16 pass over the same operands' 16-bit planes, 5/3 MMA units "
                        "per MAC of a train step: outputs and gradients within 1e-3 of the fp32 reference (same referees as c8)",
                 "c8h": "forward exactly as c8 (G output within 1e-3 of the reference); the backward GEMMs of the c8 layers run ONE fp16 "
                        "pass with a dynamic power-of-two scale on dz: gradients TF32-class (gate 5e-3 on the packed gradient, "
                        "tests/test_gpu_network.py), the accuracy class of the reference's own CUDA default (cudnn.allow_tf32)",
                 "mixed": "forward split-bf16 x3 (G output within 1e-3 of the reference), backward GEMMs single bf16 pass (gradients ~1e-2)",
                 "fast": "bf16 single pass everywhere: G output ~1e-2 rel. error vs fp32 reference (outside the 1e-3 gate)"}
    eng.set_precision(modes[args.precision])
    pkg.set_lean(bool(args.lean))

    models = ts.build_models(pkg.Generator, pkg.Discriminator, dev, seed=0)
    if args.optimizer == "fused":
        g_opt = pkg.FusedAdam(models[:2], lr=2e-4, betas=(0.5, 0.999))
        d_opt = pkg.FusedAdam(models[2:], lr=1e-4, betas=(0.5, 0.999))
    else:
        g_opt, d_opt = ts.build_optimizers(models)
    sync = pkg.GradSync([models[:2], models[2:]]) if world > 1 else None

    host = [t.pin_memory() for t in synthetic_batch_host(B, T_FRAMES, seed=1234 + rank)]
    resident = [t.to(dev, non_blocking=True) for t in host]
    h2d = sum(t.numel() * 4 for t in host)
    losses = {}

    def step_resident():
        losses["g"], losses["d"] = ts.train_step(models, g_opt, d_opt, resident)

    def step_e2e():
        batch = [t.to(dev, non_blocking=True) for t in host]   # H2D from pinned memory, every step
        g, d = ts.train_step(models, g_opt, d_opt, batch)
        losses["gh"], losses["dh"] = g.item(), d.item()        # D2H read of the step's result (train.py:302-304)

    for _ in range(W):
        step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launch_count()
    ms = timed_steps(step_resident, K, world)
    launches = eng.launch_count() - l0
    ms_e2e = timed_steps(step_e2e, K, world)
    clocks = sampler.stop() if sampler else None

    dp_attr = None
    if sync is not None:
        # where the data-parallel overhead goes: device time of the two all-reduces per step, measured with
        # CUDA events over a few extra steps (the timed region above runs without them)
        sync.enable_timing(True)
        ms_t = timed_steps(step_resident, 3, world)
        coll_ms, n_coll = sync.collective_ms()
        sync.enable_timing(False)
        dp_attr = {"steps": 3, "ms_per_step": ms_t / 3, "allreduce_ms_per_step": coll_ms / 3, "allreduces_per_step": n_coll / 3,
                   "note": "rank 0's device time inside ncclAllReduce (includes waiting for the slowest rank to arrive)"}

    frames = world * B * T_FRAMES
    value = frames * K / (ms * 1e-3)
    e2e_value = frames * K / (ms_e2e * 1e-3)

    # north_star's own target line: Generator forward + backward (parameter gradients, input not
    # requiring grad) at this batch, 58.636 algorithmic GFLOP per 80x64 sample
    peaks = load_peaks()
    G0 = models[0]
    gx, gm = resident[0], resident[1]

    def g_fwd_bwd():
        G0.zero_grad(set_to_none=True)
        G0(gx, gm).sum().backward()

    gfb = {}
    G0.train()
    for gmode in ("parity", "c8", "c8w", "c8h", "mixed", "fast"):
        eng.set_precision(modes[gmode])
        for _ in range(3):
            g_fwd_bwd()
        ms_g = timed_steps(g_fwd_bwd, 10, world) / 10
        tf = 58.636e9 * B / (ms_g * 1e-3) / 1e12
        gfb[gmode] = {"ms": ms_g, "algorithmic_tflops": tf, "frac_of_sustained_peak": tf / peaks["bf16_tflops_sustained"],
                      "frac_of_burst_peak": tf / peaks["bf16_tflops"]}
    eng.set_precision(modes[args.precision])
    G0.zero_grad(set_to_none=True)

    # roofline: per-launch CUDA-event timing of the tensor-core kernels over extra steps
    eng.set_overlap(False)      # kernels timed one at a time on their stream (no concurrent wgrad stream)
    step_resident()
    torch.cuda.synchronize()
    eng.profile_enable(True)
    for _ in range(max(args.profile_steps, 1)):
        step_resident()
    torch.cuda.synchronize()
    kinds = eng.profile_collect_kinds()
    eng.profile_enable(False)
    eng.set_overlap(True)
    peak = peaks["bf16_tflops_sustained"]
    psteps = max(args.profile_steps, 1)

    def tf(k):
        return k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0.0

    convs = [k for k in kinds if "wgrad" not in k["kernel"] and k["launches"] > 0]
    wgs = [k for k in kinds if "wgrad" in k["kernel"] and k["launches"] > 0]
    dom = max(convs, key=lambda k: k["ms"])            # the DOMINANT kernel family of the step
    conv_all = {"ms": sum(k["ms"] for k in convs), "flops": sum(k["flops"] for k in convs), "launches": sum(k["launches"] for k in convs)}
    wg = {"ms": sum(k["ms"] for k in wgs), "flops": sum(k["flops"] for k in wgs), "launches": sum(k["launches"] for k in wgs)}
    step_flops = (ts.STEP_FLOPS_LEAN_T64 if args.lean else ts.STEP_FLOPS_STRICT_T64) * B
    # roofline.traffic: DRAM bytes per launch of the dominant kernel's OWN launches, from this round's
    # `ncu --set full` capture (tools/ncu_summary.py --json; ncu cannot run inside the timed bench)
    traffic, traffic_src = None, None
    # (c8w runs c8's forward / data-gradient kernels: same launches of the dominant kernel)
    tpath = os.path.join(ROOT, "profiles", "r02_kernel_traffic_%s.json" % ("c8" if args.precision == "c8w" else args.precision))
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        for name, rec in tj.get("kernels", {}).items():
            if name.replace(" ", "").startswith(dom["kernel"].replace(" ", "")):
                traffic = rec["dram_bytes_per_launch"]
                traffic_src = "%s (%s, summarised at commit %s): %d launches of this kernel in one G+D forward+backward at batch 64" % (
                    os.path.basename(tpath), tj.get("source"), tj.get("summarised_at_commit"), rec["launches"])
    roofline = {"bound": "tensor", "kernel": dom["kernel"] + " (implicit-GEMM forward + data-gradient convolutions, tcgen05 CTA pairs)",
                "achieved": tf(dom), "peak": peak, "unit": "TFLOP/s", "frac": tf(dom) / peak,
                "peak_source": "bf16_tflops_sustained, " + peaks["source"] + " (kernel timed inside a long step)",
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_flops_per_launch": dom["flops"] / max(dom["launches"], 1),
                "avg_launch_ms": dom["ms"] / max(dom["launches"], 1),
                "launches_per_step": dom["launches"] / psteps,
                "share_of_step": (dom["ms"] / psteps) / (ms / K),
                "timing": "CUDA events around every launch on its stream (mcgvc_profile_*), %d extra steps with the weight-gradient side stream folded into the caller's stream" % psteps,
                "all_conv_kernels": {"achieved": tf(conv_all), "frac": tf(conv_all) / peak, "launches_per_step": conv_all["launches"] / psteps,
                                     "share_of_step": (conv_all["ms"] / psteps) / (ms / K)},
                "by_kernel": [{"kernel": k["kernel"], "achieved": tf(k), "launches_per_step": k["launches"] / psteps,
                               "share_of_step": (k["ms"] / psteps) / (ms / K)} for k in sorted(kinds, key=lambda k: -k["ms"]) if k["launches"] > 0],
                "wgrad_kernel": {"achieved": tf(wg), "frac": tf(wg) / peak, "launches_per_step": wg["launches"] / psteps,
                                 "share_of_step": (wg["ms"] / psteps) / (ms / K)},
                "whole_step": {"algorithmic_tflops": step_flops / (ms / K * 1e-3) / 1e12,
                               "frac": step_flops / (ms / K * 1e-3) / 1e12 / peak},
                "note": {"parity": "split-bf16 issues 3 bf16 MMAs per algorithmic MAC (ceiling 1/3 of bf16 peak); achieved counts algorithmic FLOPs once",
                         "mixed": "forward: 3 bf16 MMAs per algorithmic MAC, backward: 1; achieved counts algorithmic FLOPs once",
                         "c8h": "forward: one fp16 MMA + two e4m3 MMAs (2x rate) per MAC = 2 units, backward: 1 fp16 MMA per MAC; achieved counts algorithmic FLOPs once",
                         "c8": "one fp16 MMA + two e4m3 MMAs (2x rate) per algorithmic MAC = 2 bf16-MMA time units (ceiling 1/2 of bf16 peak); achieved counts algorithmic FLOPs once",
                         "c8w": "forward and data-gradient convolutions (this kernel): one fp16 MMA + two e4m3 MMAs (2x rate) per algorithmic MAC = 2 units (ceiling 1/2 of bf16 peak); weight gradients: 1 fp16 MMA per MAC; achieved counts algorithmic FLOPs once",
                         "fast": "single bf16 pass"}[args.precision]}

    other = None
    if args.fast_steps > 0:
        other = []
        for other_mode in ("parity", "c8", "c8w", "c8h", "mixed", "fast"):
            if other_mode == args.precision:
                continue
            eng.set_precision(modes[other_mode])
            for _ in range(2):
                step_resident()
            ms_o = timed_steps(step_resident, args.fast_steps, world)
            other.append({"precision": other_mode, "value": frames * args.fast_steps / (ms_o * 1e-3), "unit": UNIT,
                          "ms_per_step": ms_o / args.fast_steps, "note": mode_note[other_mode] + " -- reported for context only"})
        eng.set_precision(modes[args.precision])
        if not args.lean:
            # same precision, but eval-mode modules treated as frozen: skips exactly the gradients
            # train.py discards (504 instead of 670 GFLOP per sample pair); same optimisation trajectory
            pkg.set_lean(True)
            for _ in range(2):
                step_resident()
            ms_o = timed_steps(step_resident, args.fast_steps, world)
            other.append({"precision": args.precision, "lean": True, "value": frames * args.fast_steps / (ms_o * 1e-3),
                          "unit": UNIT, "ms_per_step": ms_o / args.fast_steps,
                          "note": "MCGVC_LEAN=1: gradients the reference loop computes and then discards are not computed -- reported for context only"})
            pkg.set_lean(False)

    # SURVEY 8(f) rows f1 + f2 wired into the step behind a flag (train.py owns Adam and the loss tail, so the
    # headline keeps torch.optim.Adam and the torch loss expressions): same models, engine FusedAdam on the flat
    # buffers + the one-launch-per-term loss kernels
    opt_in = None
    if args.fast_steps > 0 and args.optimizer == "torch":
        fg = pkg.FusedAdam(models[:2], lr=2e-4, betas=(0.5, 0.999))
        fd = pkg.FusedAdam(models[2:], lr=1e-4, betas=(0.5, 0.999))

        def step_opt_in():
            ts.train_step(models, fg, fd, resident, fused_losses=True)

        for _ in range(2):
            step_opt_in()
        ms_o = timed_steps(step_opt_in, args.fast_steps, world)
        opt_in = {"what": "engine FusedAdam (f1) + fused loss tail (f2) instead of torch.optim.Adam / torch loss expressions",
                  "precision": args.precision, "ms_per_step": ms_o / args.fast_steps,
                  "value": frames * args.fast_steps / (ms_o * 1e-3), "unit": UNIT}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.cpu_batch, args.cpu_steps)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"parity": "bf16x3 split (fp32 accumulate, fp32 activations/stats)",
                      "c8": "fp16 + 2x e4m3 correction (fp32 accumulate, fp32 activations/stats); stems/heads bf16x3",
                      "c8w": "fp16 + 2x e4m3 correction in forward and data-gradient GEMMs, fp16 weight-gradient GEMMs (fp32 accumulate, fp32 activations/stats); stems/heads/trunk bf16x3",
                      "c8h": "forward fp16 + 2x e4m3 correction / backward fp16 (fp32 accumulate, fp32 activations/stats); stems/heads/trunk bf16x3",
                      "mixed": "bf16x3 split forward / bf16 backward (fp32 accumulate)", "fast": "bf16"}[args.precision],
            "data": "synthetic",
            "config": make_config(B, world, args),
            "engine": {"optimizer": "torch.optim.Adam" if args.optimizer == "torch" else "engine FusedAdam",
                       "precision_mode": args.precision, "precision_note": mode_note[args.precision], "lean": bool(args.lean),
                       "grad_allreduce": "one NCCL all-reduce (sum; 1/world folded into the gradient unpack) per optimizer step on the live gradient arena" if world > 1 else "none (1 GPU)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "other_precision": other,
            "opt_in_f1_f2": opt_in,
            "generator_fwd_bwd": {"note": "north_star target line: Generator forward+backward at batch %d, 80x64 "
                                          "(58.636 algorithmic GFLOP per sample; parity mode issues 3 MMAs per MAC, "
                                          "mixed = forward x3 / backward x1)" % B, **gfb},
            "losses_last_step": {"g": float(losses["gh"]), "d": float(losses["dh"])},
        }
        if sync is not None:
            line["engine"]["allreduces"] = sync.reductions
            line["engine"]["allreduce_bytes"] = sync.reduced_bytes
            line["engine"]["data_parallel_attribution"] = dp_attr
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
