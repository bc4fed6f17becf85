"""Timing of one conv shape under the halo kernel's ablation flags (profiling experiment)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from soccernet_calibration_sportlight_b200 import ops
    C, H, W, B = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), 64
    cp = (C + 63) // 64 * 64
    x = torch.randn(B, H, W, cp, device="cuda").half()
    w = (torch.randn((C + 15) // 16 * 16, 9 * cp, device="cuda") / 30).half()
    SL = os.environ.get("W_SLICES", "0") == "1"
    rows = w.shape[0]
    if SL:
        w = w.reshape(rows, 9 * cp // 64, 64).permute(1, 0, 2).contiguous()
    bias = torch.zeros(cp, device="cuda")
    y = torch.empty_like(x)
    for res in (None, x):
        for _ in range(3):
            ops.conv2d(x, w, bias, y, ksize=3, stride=1, cout_rows=rows, relu=True, res=res, w_slices=SL)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv2d(x, w, bias, y, ksize=3, stride=1, cout_rows=rows, relu=True, res=res, w_slices=SL)
        e1.record()
        torch.cuda.synchronize()
        print(f"  ablate={os.environ.get('CAL_DEBUG_ABLATE', '0'):>2} C={C} {H}x{W} res={'yes' if res is not None else 'no '}: "
              f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
else:
    shapes = [tuple(a.split("x")) for a in sys.argv[1:]] or [("48", "135", "240"), ("96", "68", "120")]
    for shape in shapes:
        for ab in ("0", "2", "4", "6", "8", "16", "24", "28", "30"):
            env = dict(os.environ, CAL_DEBUG_ABLATE=ab)
            subprocess.run([sys.executable, __file__, "child", *shape], env=env)
