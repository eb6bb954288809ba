"""Namespace-package override of the reference's `mask_cyclegan_vc/model.py`.

The reference has no `mask_cyclegan_vc/__init__.py`, so with
    PYTHONPATH=<repo>/maskcyclegan-vc_b200/shim:<reference checkout>
`from mask_cyclegan_vc.model import Generator, Discriminator` (train.py:14, test.py:10) resolves
here while train.py / test.py themselves still come from the reference, unchanged.
"""
import importlib.util
import os
import sys

_repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
_spec = importlib.util.spec_from_file_location("mcgvc_loader", os.path.join(_repo, "mcgvc_loader.py"))
_loader = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_loader)
_pkg = _loader.load()

Generator = _pkg.Generator
Discriminator = _pkg.Discriminator
