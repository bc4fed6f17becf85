"""Does tcgen05.mma accept a SWIZZLE_128B operand whose start is shifted by whole rows?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccernet_calibration_sportlight_b200 import _lib
L = _lib.lib()
g = torch.Generator().manual_seed(0)
x = torch.randn(256, 64, generator=g).half().cuda()
w = torch.randn(64, 64, generator=g).half().cuda()
for mode in (0, 1):
    for shift in (0, 8, 1, 2, 3, 5, 7, 9, 31, 33, 34, 66, 127, 128):
        out = torch.full((128, 64), float("nan"), device="cuda")
        st = L.cal_debug_shift_mma(x.data_ptr(), w.data_ptr(), shift, mode, out.data_ptr(), None)
        torch.cuda.synchronize()
        ref = x[shift:shift + 128].float() @ w.float().t()
        err = float((out - ref).abs().max())
        print(f"mode={mode} shift={shift:3d} status={st} max_err={err:.3e} {'OK' if err < 1e-2 else 'MISMATCH'}", flush=True)
