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
    assert a._flat_grad.data_ptr() + 5 * 4 == b._flat_grad.data_ptr()   # one contiguous arena per group
    # backward pass touching a and b: both callbacks queued, they report in one after the other
    a._cb_queued = b._cb_queued = True
    a._flat_grad.copy_(_full_batch_grads(w1, Xs, Ys))
    a._cb_queued = False
    sync.module_ready(a)
    assert sync.reductions == 0          # waits for b: ONE all-reduce per optimizer group
    b._flat_grad.copy_(_full_batch_grads(w2, Xs, Ys))
    b._cb_queued = False
    sync.module_ready(b)
    assert sync.reductions == 1 and sync.reduced_bytes == 10 * 4
    ok = torch.allclose(a._flat_grad, _full_batch_grads(w1, X, Y), atol=1e-6) and \
        torch.allclose(b._flat_grad, _full_batch_grads(w2, X, Y), atol=1e-6)
    # eval-mode module: its (discarded) gradients are not reduced
    c.training = False
    c._flat_grad.fill_(float(rank + 1))
    sync.module_ready(c)
    ok = ok and sync.reductions == 1 and float(c._flat_grad[0]) == float(rank + 1)
    # partial participation: only a took part in this backward pass
    a._flat_grad.fill_(float(rank))
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
