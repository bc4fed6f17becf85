"""Where the fused BasicBlock kernel's roles wait (profiling experiment): cycles in barrier waits per role."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
dbg = torch.zeros(148 * 5 * 3 + 148 * 5, dtype=torch.int64, device="cuda")
os.environ["CAL_DEBUG_TIMELINE"] = hex(dbg.data_ptr())
from soccernet_calibration_sportlight_b200 import ops
C, H, W, B = 48, 135, 240, 64
x = torch.randn(B, H, W, 64, device="cuda").half(); x[..., C:] = 0
ws = [(torch.randn(9, 48, 64, device="cuda") / 30).half() for _ in range(2)]
bias = [torch.randn(64, device="cuda") * 0.1 for _ in range(2)]
y = torch.empty_like(x)
for _ in range(3):
    ops.basicblock(x, ws[0], bias[0], ws[1], bias[1], y, rows=48, c=C)
torch.cuda.synchronize()
raw = dbg.cpu().numpy()
d = raw[:148 * 15].reshape(148, 5, 3).astype(float)
ph = raw[148 * 15:].reshape(148, 5).astype(float)
names = ["producer  (inEmpty, -)", "MMA conv1 (inFull, t1empty)", "MMA conv2 (midFull, t2empty)", "epi conv1 (t1full, midEmpty)", "epi conv2 (t2full, -)"]
print("median over CTAs: cycles waiting on barrier A, barrier B, role total; steps per CTA ~", 4 * 34)
for r, n in enumerate(names):
    m = [float(sorted(d[:, r, k])[74]) for k in range(3)]
    print(f"  {n:30s} wait A {m[0]:9.0f}  wait B {m[1]:9.0f}  total {m[2]:9.0f}   busy {m[2] - m[0] - m[1]:9.0f}")

import numpy as np
print("epi conv1 phases (median CTA, cycles): tcgen05.wait::ld %d | fence::after + midEmpty wait %d | TMEM loads + math + smem stores (incl. wait::ld) %d | fence.proxy.async %d | syncwarp + arrive %d" % tuple(np.median(ph, axis=0)))
