"""Timing of the generic tcgen05 conv kernel on the pipeline's stride-2 / 1x1 shapes (profiling experiment)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from soccernet_calibration_sportlight_b200 import ops
B = 64
SHAPES = [  # ksize, stride, Cin, Cout, Hin, Win, residual
    (3, 2, 64, 64, 135, 240, False), (3, 2, 64, 128, 135, 240, False), (3, 2, 128, 192, 68, 120, False),
    (3, 2, 192, 384, 34, 60, False), (1, 1, 64, 256, 135, 240, True), (1, 1, 256, 64, 135, 240, False),
    (1, 1, 128, 64, 68, 120, False), (1, 1, 192, 128, 34, 60, False), (3, 1, 256, 64, 135, 240, False),
]
for ks, st, cin, cout, h, w, res in SHAPES:
    cp, op = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    ho, wo = (h - 1) // st + 1, (w - 1) // st + 1
    x = torch.randn(B, h, w, cp, device="cuda").half()
    wt = (torch.randn(op, ks * ks * cp, device="cuda") / 30).half()
    bias = torch.zeros(op, device="cuda")
    y = torch.empty(B, ho, wo, op, device="cuda", dtype=torch.half)
    r = torch.randn_like(y) if res else None
    for _ in range(3):
        ops.conv2d(x, wt, bias, y, ksize=ks, stride=st, cout_rows=op, relu=True, res=r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv2d(x, wt, bias, y, ksize=ks, stride=st, cout_rows=op, relu=True, res=r)
    e1.record(); torch.cuda.synchronize()
    print(f"  k{ks}s{st} {cin}->{cout} @{ho}x{wo} res={int(res)}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
