"""Which (LBO, SBO) assignment does tcgen05.mma expect for an MN-major SWIZZLE_128B B operand?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccernet_calibration_sportlight_b200 import _lib
L = _lib.lib()
g = torch.Generator().manual_seed(0)
x = torch.randn(128, 64, generator=g).half().cuda()
y = torch.randn(64, 128, generator=g).half().cuda()
ref = x.float() @ y.float()
for mode in (0, 1):
    out = torch.full((128, 128), float("nan"), device="cuda")
    st = L.cal_debug_mn_mma(x.data_ptr(), y.data_ptr(), mode, out.data_ptr(), None)
    torch.cuda.synchronize()
    err = float((out - ref).abs().max())
    print(f"mode={mode} (LBO,SBO)={'(8192,1024)' if mode == 0 else '(1024,8192)'} status={st} max_err={err:.3e} "
          f"{'OK' if err < 1e-2 else 'MISMATCH'}", flush=True)
