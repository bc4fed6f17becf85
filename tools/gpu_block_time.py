"""BasicBlock at the network's shape: one fused kernel vs the two conv launches (profiling experiment)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from soccernet_calibration_sportlight_b200 import ops
C, H, W, B = 48, 135, 240, int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(B, H, W, 64, device="cuda").half()
x[..., C:] = 0
ws = [(torch.randn(9, 48, 64, device="cuda") / 30).half() for _ in range(2)]
for w in ws: w[..., C:] = 0
bias = [torch.randn(64, device="cuda") * 0.1 for _ in range(2)]
t = torch.empty_like(x); y = torch.empty_like(x); y2 = torch.empty_like(x)
def two():
    ops.conv2d(x, ws[0], bias[0], t, ksize=3, stride=1, cout_rows=48, relu=True, cin=C, w_slices=True)
    ops.conv2d(t, ws[1], bias[1], y2, ksize=3, stride=1, cout_rows=48, relu=True, res=x, cin=C, w_slices=True)
def one():
    ops.basicblock(x, ws[0], bias[0], ws[1], bias[1], y, rows=48, c=C)
for name, fn in (("two launches", two), ("fused", one), ("two launches", two), ("fused", one)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:14s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per block", flush=True)
print("bit-identical:", torch.equal(y, y2))
