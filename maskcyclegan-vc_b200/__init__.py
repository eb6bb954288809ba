"""B200-native (sm_100a) conv engine for the MaskCycleGAN-VC Generator / Discriminator hot path.

The directory name contains a hyphen, so import it through `mcgvc_loader.load()` at the repo root
(registered in sys.modules as `maskcyclegan_vc_b200`), or put `shim/` on PYTHONPATH to override the
reference's `mask_cyclegan_vc.model` in place.
"""
from . import engine  # noqa: F401
from .model import Discriminator, Generator, is_lean, set_lean  # noqa: F401
from .parallel import GradSync, shard_batch  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .datafeed import DeviceVCDataFeed, draw_selection  # noqa: F401
from . import losses  # noqa: F401
