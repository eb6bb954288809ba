"""GPU, 2 ranks over NCCL (NVLink): data-parallel equivalence with the REAL engine modules, arenas and
collective: per-rank halves of a batch through GradSync must give, on every
rank, the single-process gradient of the whole batch -- an exact property because every op is
per-sample (InstanceNorm) and the loss is a batch mean.  Also covers the checkpoint bounce
(`model.to('cpu')` / `.to(device)`, saver/model_saver.py:64,74) between two synchronised steps.
Skipped when fewer than two GPUs are visible (run once via `gpurun --gpus 2`, log under profiles/)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _adv_backward(G, D, x, m):
    G.zero_grad(set_to_none=True)
    D.zero_grad(set_to_none=True)
    loss = torch.mean((1 - D(G(x, m))) ** 2)
    loss.backward()
    torch.cuda.synchronize()
    return G._flat_grad.clone(), D._flat_grad.clone()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import mcgvc_loader
    pkg = mcgvc_loader.load()
    import maskcyclegan_oracle as O
    torch.manual_seed(0)
    G, D = pkg.Generator().to(dev), pkg.Discriminator().to(dev)
    G.train(); D.train()
    x, m, _, _ = O.synthetic_batch(4, 64, seed=99)
    x, m = x.to(dev), m.to(dev)
    out = {}
    # single-process gradient of the whole batch (no GradSync yet)
    g_full, d_full = _adv_backward(G, D, x, m)
    sync = pkg.GradSync([[G], [D]])
    xs, ms = pkg.shard_batch((x, m), rank, world)
    g_dp, d_dp = _adv_backward(G, D, xs, ms)
    out["reductions"] = sync.reductions
    out["g"] = _rel(g_dp, g_full)
    out["d"] = _rel(d_dp, d_full)
    out["d_arena_floats"] = sync.groups[1]["arena"].numel()
    # every rank holds the same bits after the all-reduce
    gather = [torch.empty_like(g_dp) for _ in range(world)]
    dist.all_gather(gather, g_dp)
    out["ranks_identical"] = bool(all(torch.equal(gather[0], t) for t in gather))
    # checkpoint bounce between two synchronised steps: the arena binding must survive
    for mod in (G, D):
        mod.to("cpu")
        mod.to(dev)
    g_dp2, d_dp2 = _adv_backward(G, D, xs, ms)
    out["reductions_after_bounce"] = sync.reductions
    out["g_after_bounce"] = _rel(g_dp2, g_full)
    out["d_after_bounce"] = _rel(d_dp2, d_full)
    out["still_in_arena"] = G._flat_grad.data_ptr() == sync.groups[0]["arena"].data_ptr()
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_gradient_equals_single_process_gradient():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        o = ret[r]
        assert o["reductions"] == 2, o                 # ONE all-reduce per optimizer group
        assert o["reductions_after_bounce"] == 4, o
        assert o["d_arena_floats"] == (6202881 + 63) // 64 * 64, o   # 99.2 MB / 4: no downSample4 slot
        assert o["ranks_identical"] and o["still_in_arena"], o
        # not bit-equal: batch-4 sums vs two batch-2 sums, and in the C8 modes the power-of-two scale of
        # each dz tensor comes from a bound over the WHOLE (local) batch, so the halves round their
        # operands on a different grid than the full batch (measured 5e-5 in C8).  In the default C8W
        # mode the weight-gradient operands are additionally rounded to fp16 AFTER the two runs' ~1e-5
        # run-to-run noise, which decorrelates part of that rounding (its own size: 2.3e-4 of the
        # gradient norm): gate 5e-4.  A wrong average or a missed reduction would read as O(1).
        for k in ("g", "d", "g_after_bounce", "d_after_bounce"):
            assert o[k] < 5e-4, (r, k, o[k])
    print("2-rank NCCL vs single process:", {k: "%.2e" % ret[0][k] for k in ("g", "d", "g_after_bounce", "d_after_bounce")})
