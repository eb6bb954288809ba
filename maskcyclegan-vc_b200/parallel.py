"""Data-parallel training over the GPUs of one box: one process per GPU, torch.distributed (NCCL
over NVLink 5 / NVSwitch) for the plumbing.

The path shards over the batch (every op is per-sample: InstanceNorm, not BatchNorm; all losses are
batch means, train.py:219-232,276-288), so with equal per-rank batches the global-batch gradient is
the average of per-rank gradients.  The only exchange step is therefore ONE all-reduce per
optimizer step on the packed gradient buffer of the modules that optimizer owns:
49 075 458 floats (196.3 MB) for the two generators, 24 811 524 live floats (99.2 MB) for the four
discriminators -- the modules' gradient buffers use the "live" layout, i.e. the unused downSample4
tensors (10 488 832 floats per discriminator, model.py:316-320 vs :340-349) have no slot in them.

GradSync re-homes the modules' live gradient buffers into one contiguous arena per group; each
module reports in from its end-of-backward callback and the arena is all-reduced once, when the
last training-mode module of the group that took part in this backward pass has reported.  The
1/world average is folded into the modules' unpack pass (`unpack_scale`), so the collective is a
plain sum and no extra pass over the arena follows it.
"""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, groups, process_group=None, average=True):
        """groups: list of lists of engine modules (e.g. [[G_A2B, G_B2A], [D_A, D_B, D_A2, D_B2]])."""
        self.pg = process_group
        self.average = average
        self.groups = []
        for mods in groups:
            mods = list(mods)
            # every module's slice starts 256-byte aligned (FusedAdam and the unpack kernels use 16-byte accesses)
            sizes = [_live_count(m) for m in mods]
            offs, total = [], 0
            for n in sizes:
                offs.append(total)
                total += (n + 63) // 64 * 64
            arena = torch.zeros(total, dtype=torch.float32, device=mods[0]._flat.device)
            for m, off, n in zip(mods, offs, sizes):
                m._flat_grad = arena[off:off + n]
                m._grad_sync = self
                for p in m.parameters():
                    p.grad = None
            self.groups.append({"mods": mods, "arena": arena, "pending": set()})
        self.reductions = 0
        self.reduced_bytes = 0
        # optional attribution of the collective's cost (bench.py): CUDA events around every all-reduce
        self.timing = False
        self._events = []

    def world(self):
        return dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1

    def unpack_scale(self):
        """Factor the modules apply while writing their gradient into the arena."""
        return 1.0 / self.world() if self.average else 1.0

    def module_ready(self, module):
        """Called by a module at the end of a backward pass, after its flat gradient is final."""
        for g in self.groups:
            if module in g["mods"]:
                g["pending"].add(id(module))
                # modules of the group that will still report in this backward pass
                waiting = [m for m in g["mods"] if (m._cb_queued or m._touched) and id(m) not in g["pending"]]
                if not waiting:
                    self._reduce(g)
                return

    def _reduce(self, g):
        mods = [m for m in g["mods"] if id(m) in g["pending"]]
        g["pending"] = set()
        # gradients the training loop discards (eval-mode modules, train.py:197-200,247-248) are
        # not worth wire time: only groups with a training-mode participant are reduced
        if not any(m.training for m in mods):
            return
        if self.world() == 1:
            return
        if len(mods) == len(g["mods"]):
            buf = g["arena"]
        else:
            buf = None
        if buf is not None:
            self._allreduce(buf)
        else:
            for m in mods:
                self._allreduce(m._flat_grad)

    def enable_timing(self, on=True):
        self.timing = bool(on)
        self._events = []

    def collective_ms(self):
        """Summed device time of the all-reduces since enable_timing() (synchronises)."""
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._events)
        n = len(self._events)
        self._events = []
        return ms, n

    def _allreduce(self, buf):
        if self.timing and buf.is_cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
            b.record()
            self._events.append((a, b))
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)   # average: see unpack_scale()
        self.reductions += 1
        self.reduced_bytes += buf.numel() * 4


def _live_count(m):
    """Floats in a module's live gradient layout (engine modules), or all of them (anything else)."""
    model = getattr(m, "MODEL", None)
    if model is None:
        return m._flat.numel()
    from . import engine
    return engine.live_grad_count(model)


def shard_batch(batch, rank, world):
    """rank r gets samples [r*B_loc, (r+1)*B_loc) of each tensor in `batch` (equal shards)."""
    out = []
    for t in batch:
        B = t.shape[0]
        if B % world:
            raise ValueError("global batch %d is not divisible by world size %d" % (B, world))
        k = B // world
        out.append(t[rank * k:(rank + 1) * k])
    return tuple(out)
