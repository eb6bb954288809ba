"""Quick report of the C8 precision mode on a B200: parity of a G -> D adversarial pass against the
CPU oracle, next to the split-bf16 parity mode.    python tools/c8_check.py [B] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import net_check  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
G, D, gs, ds = net_check.build_models(0)
eng = net_check.ENG
for name, mode in (("parity (split-bf16 x3)", eng.PRECISION_PARITY), ("c8 (fp16 + 2 x e4m3)", eng.PRECISION_C8)):
    eng.set_precision(mode)
    print("== %s" % name, flush=True)
    net_check.check_backward(G, D, gs, ds, B, T, verbose=True)
eng.set_precision(eng.PRECISION_PARITY)
