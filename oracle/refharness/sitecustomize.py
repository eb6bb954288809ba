"""TEST INFRASTRUCTURE: lets the reference's UNMODIFIED train.py / test.py run in this image.

Put this directory on PYTHONPATH (Python imports `sitecustomize` at start-up).  It
  * provides stand-ins for three plotting / logging packages the image does not have (librosa,
    matplotlib, tensorboardX) -- they are only used for figures and TensorBoard output
    (mask_cyclegan_vc/utils.py:16-22,42-66, logger/base_logger.py:5);
  * makes `torch.hub.load('descriptinc/melgan-neurips', ...)` (train.py:46, test.py:36; needs the
    network) return a deterministic stand-in vocoder with the `.inverse(mel)` method the drivers call;
  * replaces `torchaudio.save` (test.py:102-103; needs a codec backend) with a float32 .npy dump;
  * records every scalar the reference's logger emits, at full precision, as JSON lines in
    $MCGVC_SCALAR_LOG (the reference's own .log file rounds losses to 3 decimals).
Nothing here touches the model, the losses, the optimizers or the data path.
"""
import io
import json
import os
import sys
import types


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    # ---- tensorboardX
    class SummaryWriter:
        def __init__(self, *a, **k):
            self._path = os.environ.get("MCGVC_SCALAR_LOG")

        def add_scalar(self, name, value, step=None, *a, **k):
            if self._path:
                with open(self._path, "a") as f:
                    f.write(json.dumps({"name": name, "value": float(value), "step": step}) + "\n")

        def __getattr__(self, name):      # add_image / add_audio / add_text / close ...
            return lambda *a, **k: None

    try:
        import tensorboardX  # noqa: F401
    except ImportError:
        _module("tensorboardX", SummaryWriter=SummaryWriter)

    # ---- librosa (power_to_db + display.specshow are only used to draw a figure)
    try:
        import librosa  # noqa: F401
    except ImportError:
        import numpy as np
        disp = _module("librosa.display", specshow=lambda *a, **k: None)
        _module("librosa", display=disp, power_to_db=lambda S, ref=1.0, **k: np.asarray(S))

    # ---- matplotlib (a figure is rendered to JPEG and read back with PIL, utils.py:54-62)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        class _Fig:
            pass

        def _savefig(buf, format="jpeg", **k):
            from PIL import Image
            Image.new("RGB", (8, 8)).save(buf, format="JPEG")

        plt = _module("matplotlib.pyplot", subplots=lambda *a, **k: (_Fig(), object()),
                      savefig=_savefig, close=lambda *a, **k: None)
        agg = _module("matplotlib.backends.backend_agg", FigureCanvasAgg=lambda fig: None)
        backends = _module("matplotlib.backends", backend_agg=agg)
        figure = _module("matplotlib.figure", Figure=_Fig)
        _module("matplotlib", pyplot=plt, backends=backends, figure=figure, use=lambda *a, **k: None)


class _StandInVocoder:
    """`.inverse(mel)`: (1, 80, T) mel -> (1, 80*T) "waveform" that is simply the mel, flattened --
    information-preserving, so comparing two arms' audio files compares their Generator outputs
    element by element.  The drivers only hand its output to the logger / audio writer."""

    def inverse(self, mel):
        import torch
        return torch.as_tensor(mel).float().reshape(1, -1)


def _patch_torch():
    import torch          # every driver imports it anyway
    torch.hub.load = lambda *a, **k: _StandInVocoder()
    try:
        import torchaudio

        def _save(path, wav, sample_rate=None, **k):
            import numpy as np
            np.save(path + ".npy", torch.as_tensor(wav).detach().cpu().float().numpy())

        torchaudio.save = _save
    except Exception:
        pass


if os.environ.get("MCGVC_REF_HARNESS", "1") == "1":
    _install_stubs()
    _patch_torch()
