"""Drop-in `Generator` / `Discriminator` for MaskCycleGAN-VC backed by the sm_100a engine.

Boundary contract (reference mask_cyclegan_vc/model.py:110,239 and :287,340; callers
train.py:103-122,203-216,255-273, test.py:66-107, saver/model_saver.py:58-74,117):

* same constructor signatures, class names, `forward(x, mask)` / `forward(x)` signatures;
* `parameters()` / `state_dict()` expose the reference's tensors (names, shapes, order, including
  the `convLayer.*` alias of `upSample2.*` and the Discriminator's unused `downSample4.*`), created
  with the same initialisers in the same order, so `torch.manual_seed(s)` gives the same weights;
* forward / backward run entirely in libmcgvc.so (hand-written CUDA); the torch modules built in
  `__init__` only own the parameters.  There is no CPU or PyTorch-arithmetic fallback: calling
  forward on CPU tensors raises.

Gradients: each module owns one flat fp32 parameter buffer and one flat gradient buffer (the
parameters and their `.grad`s are views into them).  The engine accumulates weight gradients in
its own layout during backward; an end-of-backward callback converts them once into the flat
gradient buffer and publishes `p.grad` (this is what replaces autograd's AccumulateGrad nodes, and
where data-parallel training hooks its single all-reduce, see parallel.py).
"""
import os

import torch
import torch.nn as nn
from torch.autograd import Function, Variable

from . import engine

_LEAN = os.environ.get("MCGVC_LEAN", "0") == "1"


def set_lean(flag):
    """lean=True: a module in eval() mode is treated as frozen (no weight gradients, outputs carry
    no graph unless the input needs one).  train.py toggles exactly the modules whose gradients it
    later discards (train.py:195-200,247-252), so its optimisation trajectory is unchanged, but
    `.grad` of eval-mode modules is no longer populated.  Default False = everything autograd
    would compute."""
    global _LEAN
    _LEAN = bool(flag)


def is_lean():
    return _LEAN


class _Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices equal to the reference's."""

    def forward(self, x):  # pragma: no cover - never used for arithmetic
        raise RuntimeError("placeholder module; the engine computes the whole network")


def _seq(*mods):
    return nn.Sequential(*mods)


class _EngineModule(nn.Module):
    MODEL = None

    def __init__(self):
        super().__init__()
        self._flat = None
        self._flat_grad = None
        self._gblob = None
        self._packed = None
        self._packed_version = None
        self._anchor = None
        self._cb_queued = False
        self._touched = False
        self._live = None          # parameters that receive gradients (excludes dead tensors)
        self._param_list = None    # cached list(self.parameters()) (walked on every forward)
        self._grad_sync = None     # set by parallel.GradSync
        self._weights_epoch = 0    # bumped by optim.FusedAdam (in-place update of the flat buffer)

    # ------------------------------------------------------------------ parameter storage
    def _unique_params(self):
        if self._param_list is None:
            self._param_list = list(self.parameters())
        return self._param_list

    def _reset_backward_state(self):
        self._cb_queued = False
        self._touched = False

    def invalidate_packed(self):
        """Force a re-pack of the engine-layout weights at the next forward.  Needed only after
        in-place edits that bypass the Parameters' version counters (`p.data.copy_()`,
        `m.weight.data.normal_()`, raw writes into `_flat`); optimizers, `load_state_dict` and
        `.to()` are tracked automatically."""
        self._packed = None
        self._packed_version = None

    def _flatten(self):
        """Re-home all parameters into one flat buffer (reference parameters() order)."""
        params = self._unique_params()
        if not params:
            return
        dev = params[0].device
        n = sum(p.numel() for p in params)
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = flat[off:off + k].view(p.shape)
                if p.grad is not None:
                    p.grad = None
                off += k
        self._flat = flat
        # The flat gradient buffer survives a re-flatten when it still fits (same size, and on the
        # device the parameters live on or will come back to): ModelSaver.save() bounces every
        # model through .to('cpu') / .to(device) (model_saver.py:64,74), and a GradSync arena slice
        # dropped here would silently stop gradient synchronisation after the first checkpoint.
        fg = self._flat_grad
        if fg is not None and dev.type == "cuda" and fg.device != dev:
            if self._grad_sync is not None and dev.type == "cuda":
                raise engine.EngineError(
                    "%s moved from %s to %s while registered with GradSync: build the GradSync after "
                    "the final .to(device)" % (type(self).__name__, fg.device, dev))
            self._flat_grad = None
        self._gblob = None
        self._packed = None
        self._packed_version = None
        self._anchor = None
        self._param_list = None
        self._reset_backward_state()

    def _apply(self, fn, recurse=True):
        # .to()/.cuda()/.cpu() (train.py:103-110, model_saver.py:64,74) replace p.data: re-flatten
        super()._apply(fn, recurse)
        self._flatten()
        return self

    def _ensure_device_state(self):
        if self._flat is None:
            self._flatten()
        dev = self._flat.device
        if dev.type != "cuda":
            raise engine.EngineError(
                "%s lives on %s: the B200 engine only runs on CUDA devices (no CPU fallback)"
                % (type(self).__name__, dev))
        if self._flat.numel() != engine.param_count(self.MODEL):
            raise engine.EngineError("parameter count mismatch with the engine's model table")
        if self._gblob is None:
            self._gblob = torch.zeros(engine.grad_blob_floats(self.MODEL), dtype=torch.float32, device=dev)
            if self._flat_grad is None or self._flat_grad.device != dev:
                # live layout: the flat layout minus parameters that never get a gradient
                self._flat_grad = torch.zeros(engine.live_grad_count(self.MODEL), dtype=torch.float32, device=dev)
            self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        if self._cb_queued:
            # (only forward calls come through here, and the engine never runs a forward inside a
            # backward pass.)  A previous backward pass raised before autograd ran the
            # end-of-backward callback (OOM, EngineError, precision-mode guard): its flag would
            # otherwise stop every later pass from publishing gradients, and its partial sums would
            # leak into the next publish
            self._reset_backward_state()
            self._gblob.zero_()

    def _weights_version(self):
        # the packed layout depends on the precision mode (C8 keeps fp16 + e4m3 planes)
        return (engine.pack_class(), self._weights_epoch * 1000003 + sum(p._version for p in self._unique_params()))

    def _packed_weights(self):
        self._ensure_device_state()
        v = self._weights_version()
        if self._packed is None or v != self._packed_version:
            self._packed = engine.pack_weights(self.MODEL, self._flat)
            self._packed_version = v
        return self._packed

    # ------------------------------------------------------------------ gradient publication
    def _need_wgrad(self):
        return self.training or not _LEAN

    def _queue_publish(self):
        self._touched = True
        if not self._cb_queued:
            self._cb_queued = True
            Variable._execution_engine.queue_callback(self._publish_grads)

    def _publish_grads(self):
        """End of a backward pass: engine-layout gradient blob -> flat reference-order gradient,
        `p.grad` views installed (accumulating if the caller did not zero them)."""
        self._cb_queued = False
        if not self._touched:
            return
        self._touched = False
        live = self._live_params()
        fresh = live[0].grad is None
        if fresh:
            self._flat_grad.zero_()
        # with GradSync the data-parallel average is folded into this pass (the all-reduce is a sum)
        scale = self._grad_sync.unpack_scale() if self._grad_sync is not None else 1.0
        engine.unpack_grads_live(self.MODEL, self._gblob, self._flat_grad, scale)
        self._gblob.zero_()
        if self._grad_sync is not None:
            self._grad_sync.module_ready(self)
        if fresh:
            off = 0
            for p in live:          # live parameters in parameters() order == live gradient layout
                k = p.numel()
                p.grad = self._flat_grad[off:off + k].view(p.shape)
                off += k

    def _live_params(self):
        if self._live is None:
            dead = self._dead_prefixes()
            self._live = [p for n, p in self.named_parameters()
                          if not any(n.startswith(d) for d in dead)]
        return self._live

    def _dead_prefixes(self):
        return ()


# ---------------------------------------------------------------------------------------------
def _check_mode(ctx):
    # packed weights and saved operands are laid out for the precision mode of the forward pass
    # (split-bf16 planes vs. the C8 fp16 + e4m3 planes): a backward pass in another mode would misread them
    if engine.get_precision() != ctx.precision:
        raise engine.EngineError("precision mode changed between forward (%d) and backward (%d) of one graph"
                                 % (ctx.precision, engine.get_precision()))


def _check_saved(ctx, what):
    if ctx.saved is None:
        raise engine.EngineError(
            "%s: the saved-activation blob of this forward call was already consumed; "
            "backward(retain_graph=True) followed by a second backward (and double backward) are "
            "not supported by the engine" % what)


class _GeneratorFn(Function):
    @staticmethod
    def forward(ctx, x, mask, anchor, module):
        packed = module._packed_weights()
        x = x.contiguous()
        mask = mask.contiguous()
        out, saved = engine.generator_forward(packed, x, mask)
        ctx.module = module
        ctx.packed = packed
        ctx.saved = saved
        ctx.precision = engine.get_precision()
        ctx.save_for_backward(mask)
        ctx.dims = (x.shape[0], x.shape[2])
        return out

    @staticmethod
    def backward(ctx, dout):
        module = ctx.module
        B, T = ctx.dims
        (mask,) = ctx.saved_tensors
        _check_mode(ctx)
        _check_saved(ctx, "Generator")
        need_w = module._need_wgrad()
        dx = engine.generator_backward(ctx.packed, ctx.saved, mask, dout.contiguous(), B, T,
                                       ctx.needs_input_grad[0], module._gblob if need_w else None, need_w)
        ctx.saved = None
        if need_w:
            module._queue_publish()
        return dx, None, None, None


class _DiscriminatorFn(Function):
    @staticmethod
    def forward(ctx, x, anchor, module):
        packed = module._packed_weights()
        x = x.contiguous()
        out, saved = engine.discriminator_forward(packed, x)
        ctx.module = module
        ctx.packed = packed
        ctx.saved = saved
        ctx.precision = engine.get_precision()
        ctx.save_for_backward(out)
        ctx.dims = (x.shape[0], x.shape[2])
        return out

    @staticmethod
    def backward(ctx, dout):
        module = ctx.module
        B, T = ctx.dims
        (out,) = ctx.saved_tensors
        _check_mode(ctx)
        _check_saved(ctx, "Discriminator")
        need_w = module._need_wgrad()
        dx = engine.discriminator_backward(ctx.packed, ctx.saved, out, dout.contiguous(), B, T,
                                           ctx.needs_input_grad[0], module._gblob if need_w else None, need_w)
        ctx.saved = None
        if need_w:
            module._queue_publish()
        return dx, None, None


# ---------------------------------------------------------------------------------------------
class Generator(_EngineModule):
    """Generator of MaskCycleGAN-VC (reference model.py:106-280), engine-backed."""

    MODEL = engine.GENERATOR

    def __init__(self, input_shape=(80, 64), residual_in_channels=256):
        super().__init__()
        cx = int(input_shape[0])
        r = int(residual_in_channels)
        if cx != 80 or r != 256:
            # the reference hard-codes 256 channels x 20 rows at model.py:271; nothing else runs
            raise ValueError("the engine implements the reference's only working configuration: "
                             "80 mel bins, residual_in_channels=256")
        self.flattened_channels = (cx // 4) * r
        c2 = lambda i, o, k, s, p: nn.Conv2d(i, o, k, s, p)  # noqa: E731  (parameter holders only)
        c1 = lambda i, o, k, p: nn.Conv1d(i, o, k, 1, p)     # noqa: E731
        n2 = lambda c: nn.InstanceNorm2d(c, affine=True)     # noqa: E731
        n1 = lambda c: nn.InstanceNorm1d(c, affine=True)     # noqa: E731
        # creation order == reference construction order (same RNG consumption)
        self.conv1 = c2(2, r // 2, (5, 15), 1, (2, 7))
        self.conv1_gates = c2(2, r // 2, (5, 15), 1, (2, 7))
        for name, cin in (("downSample1", r // 2), ("downSample2", r)):
            blk = nn.Module()
            blk.convLayer = _seq(c2(cin, r, 5, 2, 2), n2(r))
            blk.convLayer_gates = _seq(c2(cin, r, 5, 2, 2), n2(r))
            setattr(self, name, blk)
        self.conv2dto1dLayer = c1(self.flattened_channels, r, 1, 0)
        self.conv2dto1dLayer_tfan = n1(r)
        for i in range(1, 7):
            blk = nn.Module()
            blk.conv1d_layer = _seq(c1(r, 2 * r, 3, 1), n1(2 * r))
            blk.conv_layer_gates = _seq(c1(r, 2 * r, 3, 1), n1(2 * r))
            blk.conv1d_out_layer = _seq(c1(2 * r, r, 3, 1), n1(r))
            setattr(self, "residualLayer%d" % i, blk)
        self.conv1dto2dLayer = c1(r, self.flattened_channels, 1, 0)
        self.conv1dto2dLayer_tfan = n1(self.flattened_channels)
        # the reference's upsample() helper stores its Sequential under `self.convLayer` before
        # returning it (model.py:226-237), so `convLayer` is registered ahead of `upSample1` and
        # ends up aliasing upSample2: reproduce that registration order for state_dict parity
        up1 = _seq(c2(r, 4 * r, 5, 1, 2), _Slot(), n2(r), _Slot())
        self.convLayer = up1
        self.upSample1 = up1
        self.glu = _Slot()
        up2 = _seq(c2(r, 2 * r, 5, 1, 2), _Slot(), n2(r // 2), _Slot())
        self.convLayer = up2
        self.upSample2 = up2
        self.lastConvLayer = c2(r // 2, 1, (5, 15), 1, (2, 7))
        self._flatten()

    def forward(self, x, mask):
        self._ensure_device_state()
        if x.device != self._flat.device:
            raise engine.EngineError("input is on %s but the model is on %s" % (x.device, self._flat.device))
        x = x.float()
        mask = mask.float()
        track = torch.is_grad_enabled() and not (_LEAN and not self.training and not x.requires_grad)
        if not track:
            with torch.no_grad():
                return _GeneratorFn.apply(x, mask, self._anchor, self)
        return _GeneratorFn.apply(x, mask, self._anchor, self)


class Discriminator(_EngineModule):
    """PatchGAN discriminator (reference model.py:283-349), engine-backed."""

    MODEL = engine.DISCRIMINATOR

    def __init__(self, input_shape=(80, 64), residual_in_channels=256):
        super().__init__()
        r = int(residual_in_channels)
        if r != 256:
            raise ValueError("the engine implements residual_in_channels=256 (the reference default)")
        c2 = lambda i, o, k, s, p: nn.Conv2d(i, o, k, s, p)  # noqa: E731
        n2 = lambda c: nn.InstanceNorm2d(c, affine=True)     # noqa: E731
        self.convLayer1 = _seq(c2(1, r // 2, (3, 3), (1, 1), (1, 1)), _Slot())
        self.downSample1 = _seq(c2(r // 2, r, (3, 3), (2, 2), 1), n2(r), _Slot())
        self.downSample2 = _seq(c2(r, 2 * r, (3, 3), (2, 2), 1), n2(2 * r), _Slot())
        self.downSample3 = _seq(c2(2 * r, 4 * r, (3, 3), (2, 2), 1), n2(4 * r), _Slot())
        # constructed but never applied by the reference forward (model.py:316-320 vs :340-349):
        # kept for parameters()/state_dict()/optimizer-state compatibility, never receives a grad
        self.downSample4 = _seq(c2(4 * r, 4 * r, (1, 10), (1, 1), (0, 2)), n2(4 * r), _Slot())
        self.outputConvLayer = _seq(c2(4 * r, 1, (1, 3), (1, 1), (0, 1)))
        self._flatten()

    def _dead_prefixes(self):
        return ("downSample4.",)

    def forward(self, x):
        self._ensure_device_state()
        if x.device != self._flat.device:
            raise engine.EngineError("input is on %s but the model is on %s" % (x.device, self._flat.device))
        x = x.float()
        track = torch.is_grad_enabled() and not (_LEAN and not self.training and not x.requires_grad)
        if not track:
            with torch.no_grad():
                return _DiscriminatorFn.apply(x, self._anchor, self)
        return _DiscriminatorFn.apply(x, self._anchor, self)
