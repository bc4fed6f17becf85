"""The dominant conv launches of the step, one NVTX range ('cap') around a single warm launch of each,
so that one `ncu --set full --nvtx --nvtx-include "cap/"` run captures exactly these:
  BasicBlock 48 @135x240 in one kernel (basicblock_kernel<3,0>),
  3x3 48->48 @135x240 without / with residual (conv3x3_halo_kernel<1,2> / <1,1>),
  3x3 96->96 @68x120, 192->192 @34x60, 384->384 @17x30 (conv3x3_pair_kernel), 1x1 64->256 @135x240 (conv_tc).
python tools/ncu_shapes.py [B] [first n shapes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from soccernet_calibration_sportlight_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
LIMIT = int(sys.argv[2]) if len(sys.argv) > 2 else 99
SHAPES = [  # ksize, Cin, Cout, H, W, residual
    (3, 48, 48, 135, 240, False), (3, 48, 48, 135, 240, True), (3, 96, 96, 68, 120, False), (3, 96, 96, 68, 120, True),
    (3, 192, 192, 34, 60, False), (3, 384, 384, 17, 30, False), (1, 64, 256, 135, 240, True),
]
# the fused BasicBlock of the full-resolution branch
x = torch.randn(B, 135, 240, 64, device="cuda").half()
x[..., 48:] = 0
ws = [(torch.randn(9, 48, 64, device="cuda") / 30).half() for _ in range(2)]
bs = [torch.randn(64, device="cuda") * 0.1 for _ in range(2)]
y = torch.empty_like(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    if it == 3:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("cap")
        e0.record()
    ops.basicblock(x, ws[0], bs[0], ws[1], bs[1], y, rows=48, c=48)
e1.record()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(f"basicblock 48 @135x240: {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
del x, y
for ks, cin, cout, h, w, res in SHAPES[:LIMIT]:
    cp, op = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    rows = (cout + 15) // 16 * 16
    x = torch.randn(B, h, w, cp, device="cuda").half()
    x[..., cin:] = 0
    wt = (torch.randn(rows, ks * ks * cp, device="cuda") / 30).half()
    if ks == 3:       # slice-major, as the engine packs the 3x3 stride-1 weights (the CTA-pair kernel takes only these)
        wt = wt.reshape(rows, ks * ks * cp // 64, 64).permute(1, 0, 2).contiguous()
    bias = torch.zeros(op, device="cuda")
    y = torch.empty(B, h, w, op, device="cuda", dtype=torch.half)
    r = torch.randn_like(y) if res else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(4):
        if it == 3:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push("cap")
            e0.record()
        ops.conv2d(x, wt, bias, y, ksize=ks, stride=1, cout_rows=rows, relu=True, res=r, cin=cin, w_slices=(ks == 3))
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print(f"k{ks} {cin}->{cout} @{h}x{w} res={int(res)}: {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
