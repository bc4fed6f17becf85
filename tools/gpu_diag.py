"""GPU bring-up diagnostics: runs every kernel on small cases against torch / the
oracle, never stops at the first failure, and writes a detailed log to
gpurun_out/diag.log (only the tail of stdout comes back from a gpurun call).
    python tools/gpu_diag.py [section ...]     sections: probe conv combine stem decode
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccernet_calibration_sportlight_b200 import ops, packing  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "diag.log"), "a")
RESULTS = {}


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def section(name):
    def deco(fn):
        fn._section = name
        return fn
    return deco


dev = torch.device("cuda:0")


# ---------------------------------------------------------------------------- TMA probe
def expected_box(x, box_w, box_h, es, c0, x0, y0, n0):
    """Expected 128B-swizzled smem image: row r = iy*nx + ix holds 64 fp16 of pixel
    (y0+iy*es, x0+ix*es), channels c0..c0+63; 16-byte chunk j of row r is stored at
    chunk position j ^ (r % 8)."""
    B, H, W, Cc = x.shape
    nx, ny = (box_w + es - 1) // es, (box_h + es - 1) // es
    img = np.zeros((128, 64), dtype=np.float16)
    xn = x.cpu().numpy()
    for iy in range(ny):
        for ix in range(nx):
            yy, xx = y0 + iy * es, x0 + ix * es
            if 0 <= yy < H and 0 <= xx < W and 0 <= n0 < B:
                img[iy * nx + ix] = xn[n0, yy, xx, c0:c0 + 64]
    sw = np.zeros_like(img).reshape(128, 8, 8)
    im = img.reshape(128, 8, 8)
    for r in range(128):
        for j in range(8):
            sw[r, j ^ (r % 8)] = im[r, j]
    return sw.reshape(128, 64), nx * ny


@section("probe")
def run_probe():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 11, 13, 128, generator=g).to(torch.float16).to(dev)
    cases = [(4, 3, 1, 0, 0, 0, 0), (4, 3, 1, 64, 2, 1, 1), (8, 4, 1, 0, -1, -1, 0),
             (8, 4, 1, 64, 9, 9, 1), (16, 8, 1, 0, -1, 5, 0),
             (8, 6, 2, 0, 0, 0, 0), (8, 6, 2, 0, -1, -1, 1), (8, 6, 2, 64, 5, 7, 0), (14, 4, 2, 0, 1, 1, 0)]
    for (bw, bh, es, c0, x0, y0, n0) in cases:
        name = f"probe bw{bw} bh{bh} es{es} c{c0} x{x0} y{y0} n{n0}"
        try:
            raw = ops.tma_probe(x, bw, bh, es, c0, x0, y0, n0)
            torch.cuda.synchronize()
            got = raw.cpu().numpy().view(np.float16).reshape(128, 64)
            exp, nrows = expected_box(x, bw, bh, es, c0, x0, y0, n0)
            ok = np.array_equal(got[:nrows].view(np.uint16), exp[:nrows].view(np.uint16))
            RESULTS[name] = bool(ok)
            log(("PASS " if ok else "FAIL ") + name)
            if not ok:
                bad = np.argwhere(got[:nrows].view(np.uint16) != exp[:nrows].view(np.uint16))
                log("   first mismatches (row, col):", bad[:10].tolist(), "n_bad", len(bad))
                log("   got row0[:8]", got[0, :8], "exp", exp[0, :8])
                log("   got row1[:8]", got[1, :8], "exp", exp[1, :8])
                # is it the unswizzled image?
                un = np.zeros_like(exp)
                e3 = exp.reshape(128, 8, 8)
                for r in range(128):
                    for j in range(8):
                        un.reshape(128, 8, 8)[r, j] = e3[r, j ^ (r % 8)]
                log("   equals unswizzled:", np.array_equal(got[:nrows].view(np.uint16), un[:nrows].view(np.uint16)))
        except Exception as e:  # noqa: BLE001
            RESULTS[name] = False
            log("ERROR " + name, repr(e))


# ---------------------------------------------------------------------------- conv
def conv_case(B, H, W, ci, co, k, s, relu=True, res=False, mode=0, ncls=0, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) * (scale / (ci * k * k) ** 0.5)
    b = torch.randn(co, generator=g) * 0.1
    xh = packing.to_nhwc16(x.to(dev))
    wp, bp, rows = packing.pack_conv(w.double(), b.double())
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    cop = packing.pad_to(co)
    # fp32 reference on the fp16-rounded operands (isolates accumulation order)
    xr = packing.from_nhwc16(xh, ci)
    wr = w.to(torch.float16).float().to(dev)
    ref = F.conv2d(xr, wr, b.to(dev), stride=s, padding=k // 2)
    r16 = None
    if res:
        r = torch.randn(B, co, Ho, Wo, generator=g)
        r16 = packing.to_nhwc16(r.to(dev))
        ref = ref + packing.from_nhwc16(r16, co)
    if mode == 0:
        if relu:
            ref = F.relu(ref)
        y = torch.full((B, Ho, Wo, cop), float("nan"), dtype=torch.float16, device=dev)
        ops.conv2d(xh, wp.to(dev), bp.to(dev), y, ksize=k, stride=s, cout_rows=rows, relu=relu, res=r16)
        torch.cuda.synchronize()
        got = packing.from_nhwc16(y, co)
        padz = bool((y[..., co:] == 0).all().item()) if cop > co else True
        tol = 2e-3 * max(1.0, float(ref.abs().max())) + 1e-3
    else:
        ref = F.log_softmax(ref, 1) if mode == 1 else F.softmax(ref, 1)
        y = torch.full((B, ncls, Ho, Wo), float("nan"), dtype=torch.float32, device=dev)
        ops.conv2d(xh, wp.to(dev), bp.to(dev), y, ksize=k, stride=s, cout_rows=rows, relu=False, mode=mode,
                   n_classes=ncls)
        torch.cuda.synchronize()
        got = y
        padz = True
        tol = 2e-3
    err = (got - ref).abs()
    finite = bool(torch.isfinite(got).all().item())
    mx = float(err.max()) if finite else float("inf")
    ok = finite and mx <= tol and padz
    info = dict(max_err=mx, tol=tol, pad_zero=padz, finite=finite)
    if not ok and finite:
        e = err.max(dim=1)[0]                      # (B,Ho,Wo)
        idx = torch.nonzero(e > tol)
        info["n_bad_pixels"] = int(idx.shape[0])
        info["first_bad"] = idx[:6].tolist()
        ec = err.amax(dim=(0, 2, 3))
        info["bad_channels"] = torch.nonzero(ec > tol).flatten()[:16].tolist()
        info["err_by_rowmod8"] = [float(err[:, :, :, i::8].max()) for i in range(min(8, Wo))]
    return ok, info


@section("conv")
def run_conv():
    cases = [
        dict(B=1, H=8, W=16, ci=64, co=64, k=1, s=1),
        dict(B=1, H=8, W=16, ci=64, co=64, k=3, s=1),
        dict(B=2, H=9, W=20, ci=64, co=64, k=3, s=1, res=True),
        dict(B=2, H=17, W=30, ci=48, co=48, k=3, s=1, res=True),
        dict(B=2, H=17, W=30, ci=96, co=96, k=3, s=1, relu=False),
        dict(B=1, H=34, W=60, ci=192, co=192, k=3, s=1, res=True),
        dict(B=2, H=17, W=30, ci=384, co=384, k=3, s=1, res=True),
        dict(B=1, H=20, W=24, ci=64, co=256, k=1, s=1, relu=False),
        dict(B=1, H=20, W=24, ci=256, co=64, k=1, s=1),
        dict(B=1, H=20, W=24, ci=256, co=48, k=3, s=1),
        dict(B=2, H=16, W=24, ci=48, co=96, k=3, s=2),
        dict(B=2, H=17, W=31, ci=48, co=96, k=3, s=2),
        dict(B=1, H=135, W=240, ci=256, co=96, k=3, s=2),
        dict(B=1, H=34, W=60, ci=192, co=384, k=3, s=2, res=True, relu=True),
        dict(B=1, H=20, W=24, ci=384, co=48, k=1, s=1, relu=False),
        dict(B=1, H=20, W=24, ci=64, co=784, k=1, s=1, res=True),
        dict(B=1, H=20, W=24, ci=784, co=58, k=1, s=1, mode=1, ncls=58, scale=4.0),
        dict(B=1, H=20, W=24, ci=720, co=23, k=1, s=1, mode=2, ncls=23, scale=4.0),
        dict(B=3, H=135, W=240, ci=48, co=48, k=3, s=1, res=True),
        dict(B=5, H=68, W=120, ci=96, co=96, k=3, s=1, res=True),
    ]
    for i, c in enumerate(cases):
        name = "conv " + " ".join(f"{k}={v}" for k, v in c.items())
        try:
            ok, info = conv_case(seed=i, **c)
            RESULTS[name] = bool(ok)
            log(("PASS " if ok else "FAIL ") + name, json.dumps(info))
        except Exception as e:  # noqa: BLE001
            RESULTS[name] = False
            log("ERROR " + name, repr(e))
            if "CUDA" in repr(e) or "cuda" in repr(e):
                log(traceback.format_exc())
                break


# ---------------------------------------------------------------------------- combine
@section("combine")
def run_combine():
    g = torch.Generator().manual_seed(3)
    for (B, H, W, Cc, srcs, relu, bias) in [
        (2, 17, 30, 48, [(17, 30), (9, 15), (5, 8)], True, False),
        (1, 135, 240, 48, [(135, 240), (68, 120), (34, 60), (17, 30)], True, False),
        (1, 68, 120, 96, [(68, 120), (68, 120), (34, 60), (17, 30)], True, False),
        (1, 54, 96, 784, [(27, 48), (14, 24), (7, 12), (4, 6)], False, True),
    ]:
        name = f"combine B{B} {H}x{W} C{Cc} srcs{srcs} relu{relu} bias{bias}"
        try:
            cp = packing.pad_to(Cc)
            ts = [torch.randn(B, Cc, h, w, generator=g) for (h, w) in srcs]
            t16 = [packing.to_nhwc16(t.to(dev)) for t in ts]
            bvec = (torch.randn(cp, generator=g) * 0.3).to(dev) if bias else None
            ref = torch.zeros(B, Cc, H, W, device=dev)
            if bias:
                ref = ref + bvec[:Cc].view(1, -1, 1, 1)
            for t in t16:
                tf = packing.from_nhwc16(t, Cc)
                if tf.shape[-2:] != (H, W):
                    tf = F.interpolate(tf, size=(H, W), mode="bilinear", align_corners=True)
                ref = ref + tf
            if relu:
                ref = F.relu(ref)
            y = torch.full((B, H, W, cp), float("nan"), dtype=torch.float16, device=dev)
            ops.fuse_combine(y, t16, bvec, relu)
            torch.cuda.synchronize()
            got = packing.from_nhwc16(y, Cc)
            mx = float((got - ref).abs().max())
            tol = 2e-3 * max(1.0, float(ref.abs().max()))
            ok = mx <= tol
            RESULTS[name] = bool(ok)
            log(("PASS " if ok else "FAIL ") + name, f"max_err={mx:.3e} tol={tol:.3e}")
        except Exception as e:  # noqa: BLE001
            RESULTS[name] = False
            log("ERROR " + name, repr(e))


# ---------------------------------------------------------------------------- stem
@section("stem")
def run_stem():
    g = torch.Generator().manual_seed(4)
    for (B, H, W) in [(1, 20, 24), (2, 21, 37), (1, 540, 960)]:
        name = f"stem B{B} {H}x{W}"
        try:
            x = torch.rand(B, 3, H, W, generator=g).to(dev)
            w = (torch.randn(64, 3, 3, 3, generator=g) * 0.3).to(dev)
            b = (torch.randn(64, generator=g) * 0.1).to(dev)
            ref = F.relu(F.conv2d(x, w, b, stride=2, padding=1))
            Ho, Wo = ref.shape[-2:]
            y = torch.full((B, Ho, Wo, 64), float("nan"), dtype=torch.float16, device=dev)
            ops.stem_conv(x, w.reshape(64, 27).contiguous(), b, y)
            torch.cuda.synchronize()
            got = packing.from_nhwc16(y, 64)
            mx = float((got - ref).abs().max())
            tol = 2e-3 * max(1.0, float(ref.abs().max()))
            ok = mx <= tol
            RESULTS[name] = bool(ok)
            log(("PASS " if ok else "FAIL ") + name, f"max_err={mx:.3e} tol={tol:.3e}")
        except Exception as e:  # noqa: BLE001
            RESULTS[name] = False
            log("ERROR " + name, repr(e))


# ---------------------------------------------------------------------------- decode
@section("decode")
def run_decode():
    from oracle import decode_ref as O
    from tests import inputs as I
    for name, logp, size in [
        ("hashed_small", I.hashed_logp(1, (2, 58, 20, 24)), (40, 48)),
        ("hashed_ragged", I.hashed_logp(2, (1, 5, 7, 13)), (21, 39)),
        ("gauss_small", I.gaussian_logp(4, 2, 58, 20, 24), (40, 48)),
        ("hashed_full", I.hashed_logp(3, (1, 58, 270, 480)), (540, 960)),
        ("gauss_full", I.gaussian_logp(5, 1, 58, 270, 480), (540, 960)),
        ("gauss_720p", I.gaussian_logp(6, 1, 58, 360, 640), (540, 960)),
    ]:
        try:
            exp = O.keypoint_decode_np(logp, size)
            got = ops.kp_decode(torch.from_numpy(logp).to(dev), size).cpu().numpy()
            ok = np.array_equal(got.view(np.uint32), exp.view(np.uint32))
            RESULTS["kp_decode " + name] = bool(ok)
            log(("PASS " if ok else "FAIL ") + "kp_decode " + name,
                "" if ok else f"n_bad={int((got != exp).sum())} first={np.argwhere(got != exp)[:5].tolist()}")
        except Exception as e:  # noqa: BLE001
            RESULTS["kp_decode " + name] = False
            log("ERROR kp_decode " + name, repr(e))
    for name, heat, sigma in [
        ("tent_small_s3", I.two_peak_heat(7, 2, 23, 17, 30), 3.0),
        ("tent_small_s6", I.two_peak_heat(8, 1, 23, 17, 30), 6.0),
        ("tent_ragged", I.two_peak_heat(10, 1, 3, 5, 11), 2.5),
        ("tent_full_s3", I.two_peak_heat(9, 1, 23, 135, 240), 3.0),
    ]:
        try:
            exp = O.line_decode_np(heat, sigma)
            got = ops.line_decode(torch.from_numpy(heat).to(dev), sigma).cpu().numpy()
            ok = np.array_equal(got.view(np.uint32), exp.view(np.uint32))
            RESULTS["line_decode " + name] = bool(ok)
            log(("PASS " if ok else "FAIL ") + "line_decode " + name,
                "" if ok else f"n_bad={int((got != exp).sum())} first={np.argwhere(got != exp)[:5].tolist()} "
                              f"maxabs={np.abs(got - exp).max()}")
        except Exception as e:  # noqa: BLE001
            RESULTS["line_decode " + name] = False
            log("ERROR line_decode " + name, repr(e))


def main():
    want = set(sys.argv[1:])
    log(f"=== gpu_diag {time.strftime('%F %T')} device={torch.cuda.get_device_name(0)} ===")
    for fn in (run_decode, run_probe, run_stem, run_combine, run_conv):
        if want and fn._section not in want:
            continue
        try:
            fn()
        except Exception:  # noqa: BLE001
            log("SECTION CRASH", fn._section, traceback.format_exc())
    npass = sum(RESULTS.values())
    log(f"=== {npass}/{len(RESULTS)} passed ===")
    with open(os.path.join(OUT, "diag.json"), "w") as f:
        json.dump(RESULTS, f, indent=1)


if __name__ == "__main__":
    main()
