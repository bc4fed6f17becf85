"""tcgen05.mma issue rate (cycles per M=128,K=16 instruction) vs N / A-operand row shift / B major-ness."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccernet_calibration_sportlight_b200 import _lib
L = _lib.lib()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
iters = 2000
for bmn in (0, 1):
    for N in (48, 64, 96, 128, 192, 256):
        row = []
        for shift in (0, 1, 8, 33):
            st = L.cal_debug_mma_rate(N, shift, iters, bmn, out.data_ptr(), None)
            torch.cuda.synchronize()
            row.append(f"shift{shift}: {int(out.item()) / iters:6.1f}")
        print(f"B {'MN' if bmn else 'K '}-major N={N:3d}  cycles/MMA  " + "  ".join(row) + f"   (tensor floor N/2 = {N / 2:.0f})", flush=True)

# the same train with a tcgen05.commit every k groups of 4 MMAs (the streamed-weight convs commit per
# weight slice = per group), optionally with the mbarrier wait + fence the real loop has, and the
# commits alone
out = torch.zeros(148, dtype=torch.int64, device="cuda")
for N in (96, 192):
    for name, wait, extra in (("none", 0, 0), ("try_wait", 1, 0), ("try_wait+fence", 1, 8), ("test_wait", 1, 16), ("smem flag", 1, 32), ("wait, commit", 1, 64)):
        for nomma in (0, 1):
            row = []
            for k in (0, 1, 2, 4, 8):
                flags = (k << 8) | (wait << 1) | extra | (nomma << 2) | (1 << 16)
                L.cal_debug_mma_rate(N, 1, iters, flags, out.data_ptr(), None)
                torch.cuda.synchronize()
                row.append(f"k={k}: {int(out[0].item()) / (iters / 4):6.1f}")
            print(f"N={N:3d} wait={name:14s} mma={1 - nomma}  cycles per group of 4 MMAs, commit every k groups  "
                  + "  ".join(row), flush=True)
