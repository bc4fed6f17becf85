"""Cycles per tile of the 3x3 kernels' tcgen05.mma trains in isolation (csrc/probe.cu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccernet_calibration_sportlight_b200 import _lib
L = _lib.lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 500
names = {0: "DX (2n | n per row)", 1: "DX, fixed A address", 2: "nine taps N=n", 3: "three taps N=3n", 4: "N=2n throughout",
         5: "N=256", 6: "DX, 2n trains first"}
def run(pat, n, nk, flags, ctas=1):
    out.zero_()
    rc = L.cal_debug_mma_pattern(pat, n, nk, iters, flags, ctas, out.data_ptr(), None)
    assert rc == 0, L.cal_last_error()
    torch.cuda.synchronize()
    return float(out[:ctas].max().item()) / iters
for n, nk in ((48, 3), (64, 4), (96, 4)):
    for pat in (0, 1, 6, 2, 3, 4, 5):
        if pat == 3 and 3 * n > 256: continue
        row = []
        for label, flags in (("zeros 1acc", 0x100), ("rand 1acc", 0x101), ("rand 2acc", 0x201), ("rand 2acc commits", 0x203),
                             ("rand 2acc commits tmem-readers", 0x207), ("same, 148 CTAs", 0x207),
                             ("commits+waits+fence", 0x20b),
                             ("DUAL rand", 0x141), ("DUAL commits+waits+fence", 0x14b), ("DUAL commits+waits+fence+readers 148", 0x14f)):
            ctas = 148 if "148" in label else 1
            row.append(f"{label}: {run(pat, n, nk, flags, ctas):7.1f}")
        nm = 9 * nk if pat == 2 else (3 * nk if pat in (3, 5) else 6 * nk)
        print(f"n={n} nk={nk} {names[pat]:24s} ({nm:2d} MMAs)  cycles/tile  " + "  ".join(row), flush=True)
