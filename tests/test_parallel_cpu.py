"""CPU, world_size 2 over gloo: the data-parallel host logic (batch sharding, one all-reduce per
optimizer group on the packed gradient arena, averaged == single-process gradient on the
concatenated batch)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakeModule:
    """Stands in for an engine module: GradSync only touches these attributes."""

    def __init__(self, n):
        self._flat = torch.zeros(n)
        self._flat_grad = torch.zeros(n)
        self._grad_sync = None
        self._cb_queued = False
        self._touched = False
        self.training = True

    def parameters(self):
        return []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _full_batch_grads(w, X, Y):
    w = w.clone().requires_grad_(True)
    loss = torch.mean((X @ w - Y) ** 2)   # batch-mean loss, as every loss in train.py:219-232
    loss.backward()
    return w.grad


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mcgvc_loader
    pkg = mcgvc_loader.load()
    g = torch.Generator().manual_seed(0)
    X = torch.randn(8, 5, generator=g)
    Y = torch.randn(8, generator=g)
    w1, w2 = torch.randn(5, generator=g), torch.randn(5, generator=g)
    Xs, Ys = pkg.shard_batch((X, Y), rank, world)
    assert Xs.shape[0] == 4 and torch.equal(Xs, X[rank * 4:(rank + 1) * 4])
    a, b, c = FakeModule(5), FakeModule(5), FakeModule(7)
    sync = pkg.GradSync([[a, b], [c]])
    # one contiguous arena per group, every module's slice 256-byte aligned
    assert a._flat_grad.data_ptr() + 64 * 4 == b._flat_grad.data_ptr()
    assert sync.unpack_scale() == 0.5    # the average is folded into the modules' unpack pass
    # backward pass touching a and b: both callbacks queued, they report in one after the other
    a._cb_queued = b._cb_queued = True
    a._flat_grad.copy_(_full_batch_grads(w1, Xs, Ys) * sync.unpack_scale())
    a._cb_queued = False
    sync.module_ready(a)
    assert sync.reductions == 0          # waits for b: ONE all-reduce per optimizer group
    b._flat_grad.copy_(_full_batch_grads(w2, Xs, Ys) * sync.unpack_scale())
    b._cb_queued = False
    sync.module_ready(b)
    assert sync.reductions == 1 and sync.reduced_bytes == 128 * 4
    ok = torch.allclose(a._flat_grad, _full_batch_grads(w1, X, Y), atol=1e-6) and \
        torch.allclose(b._flat_grad, _full_batch_grads(w2, X, Y), atol=1e-6)
    # eval-mode module: its (discarded) gradients are not reduced
    c.training = False
    c._flat_grad.fill_(float(rank + 1))
    sync.module_ready(c)
    ok = ok and sync.reductions == 1 and float(c._flat_grad[0]) == float(rank + 1)
    # partial participation: only a took part in this backward pass
    a._flat_grad.fill_(float(rank) * sync.unpack_scale())
    sync.module_ready(a)
    ok = ok and sync.reductions == 2 and abs(float(a._flat_grad[0]) - 0.5) < 1e-6
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_batch_rejects_ragged():
    import mcgvc_loader
    import pytest
    pkg = mcgvc_loader.load()
    with pytest.raises(ValueError):
        pkg.shard_batch((torch.zeros(5, 2),), 0, 2)


def test_gradsync_arena_binding_survives_checkpoint_bounce():
    """ModelSaver.save() moves every model to the CPU and back (saver/model_saver.py:64,74).  The
    re-flatten that follows must keep the module's gradient buffer inside the GradSync arena, or
    gradients silently stop being synchronised after the first checkpoint.  Host logic only (no
    compute): real engine modules on the CPU."""
    import mcgvc_loader
    pkg = mcgvc_loader.load()
    torch.manual_seed(0)
    D1, D2 = pkg.Discriminator(), pkg.Discriminator()
    sync = pkg.GradSync([[D1, D2]])
    arena = sync.groups[0]["arena"]
    live = pkg.engine.live_grad_count(pkg.engine.DISCRIMINATOR)
    assert live == 6202881                                     # downSample4 has no slot
    assert arena.numel() == 2 * ((live + 63) // 64 * 64)       # 2 x 24.8 MB instead of 2 x 66.8 MB
    ptrs = (D1._flat_grad.data_ptr(), D2._flat_grad.data_ptr())
    assert ptrs[0] == arena.data_ptr() and ptrs[1] == arena.data_ptr() + ((live + 63) // 64 * 64) * 4
    sd = {k: v.clone() for k, v in D1.state_dict().items()}
    for m in (D1, D2):
        assert m.to("cpu") is m and m.to(torch.device("cpu")) is m
    assert (D1._flat_grad.data_ptr(), D2._flat_grad.data_ptr()) == ptrs
    assert D1._grad_sync is sync and D2._grad_sync is sync
    for k, v in D1.state_dict().items():
        assert torch.equal(v, sd[k]), k
