"""The HRNet building kernels through the C ABI against plain PyTorch fp32 references of
the same op on the fp16-rounded operands (isolates accumulation order; fp32 accumulate on
tcgen05).  Tolerance: 2e-3 * max(1, |ref|_max) + 1e-3 absolute - fp16 output rounding
(2^-11 relative) plus summation-order noise; (log)softmax outputs 2e-3 absolute."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from soccernet_calibration_sportlight_b200 import ops, packing

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def conv_case(B, H, W, ci, co, k, s, relu=True, res=False, mode=0, ncls=0, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) * (scale / (ci * k * k) ** 0.5)
    b = torch.randn(co, generator=g) * 0.1
    xh = packing.to_nhwc16(x.to(dev))
    wp, bp, rows = packing.pack_conv(w.double(), b.double())
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    cop = packing.pad_to(co)
    xr = packing.from_nhwc16(xh, ci)
    wr = w.to(torch.float16).float().to(dev)
    ref = F.conv2d(xr, wr, b.to(dev), stride=s, padding=k // 2)
    r16 = None
    if res:
        r16 = packing.to_nhwc16(torch.randn(B, co, Ho, Wo, generator=g).to(dev))
        ref = ref + packing.from_nhwc16(r16, co)
    if mode == 0:
        ref = F.relu(ref) if relu else ref
        y = torch.full((B, Ho, Wo, cop), float("nan"), dtype=torch.float16, device=dev)
        ops.conv2d(xh, wp.to(dev), bp.to(dev), y, ksize=k, stride=s, cout_rows=rows, relu=relu, res=r16, cin=ci)
        got = packing.from_nhwc16(y, co)
        assert bool((y[..., co:] == 0).all()), "channel padding lanes must be written as zero"
        if k == 3:
            # slice-major weights (every (tap, chunk) slice contiguous; the halo / CTA-pair kernels) must give the same bits
            ws = wp.reshape(rows, -1, 64).permute(1, 0, 2).contiguous().to(dev)
            y2 = torch.full_like(y, float("nan"))
            try:
                ops.conv2d(xh, ws, bp.to(dev), y2, ksize=k, stride=s, cout_rows=rows, relu=relu, res=r16, cin=ci,
                           w_slices=True)
            except ops._lib.CalError as e:
                assert s == 2 and "generic kernel" in str(e)      # a stride-2 shape only the K-major kernel serves
            else:
                if s == 2 and ci > 64:
                    # the pair kernel accumulates chunk-major, the generic one tap-major: fp32 rounding, an fp16 ulp here and there
                    assert float((y.float() - y2.float()).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))
                    assert bool((y2[..., co:] == 0).all())
                else:
                    assert torch.equal(y, y2)
        tol = 2e-3 * max(1.0, float(ref.abs().max())) + 1e-3
    else:
        ref = F.log_softmax(ref, 1) if mode == 1 else F.softmax(ref, 1)
        got = torch.full((B, ncls, Ho, Wo), float("nan"), dtype=torch.float32, device=dev)
        ops.conv2d(xh, wp.to(dev), bp.to(dev), got, ksize=k, stride=s, cout_rows=rows, relu=False, mode=mode,
                   n_classes=ncls)
        tol = 2e-3
    assert bool(torch.isfinite(got).all())
    err = float((got - ref).abs().max())
    assert err <= tol, f"max err {err:.3e} > tol {tol:.3e}"


CONV_CASES = [
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
    dict(B=3, H=135, W=240, ci=48, co=48, k=3, s=2, res=True),
    dict(B=2, H=67, W=121, ci=96, co=192, k=3, s=2),
    dict(B=40, H=34, W=60, ci=48, co=96, k=3, s=2, res=True),
    dict(B=1, H=20, W=24, ci=384, co=48, k=1, s=1, relu=False),
    dict(B=1, H=20, W=24, ci=64, co=784, k=1, s=1, res=True),
    dict(B=1, H=20, W=24, ci=784, co=58, k=1, s=1, mode=1, ncls=58, scale=4.0),
    dict(B=1, H=20, W=24, ci=720, co=23, k=1, s=1, mode=2, ncls=23, scale=4.0),
    dict(B=3, H=135, W=240, ci=48, co=48, k=3, s=1, res=True),
    dict(B=5, H=68, W=120, ci=96, co=96, k=3, s=1, res=True),
    dict(B=1, H=1, W=1, ci=64, co=64, k=3, s=1),
    dict(B=1, H=3, W=130, ci=48, co=48, k=3, s=1, res=True),
    # persistent grids of 148 CTAs in clusters of 4 with multicast weight slices
    dict(B=16, H=68, W=120, ci=96, co=96, k=3, s=1, res=True),
    dict(B=8, H=34, W=60, ci=192, co=192, k=3, s=1, res=True),
    dict(B=16, H=17, W=30, ci=384, co=384, k=3, s=1, res=True),     # two N tiles: N-slowest tile order
    dict(B=9, H=34, W=60, ci=192, co=192, k=3, s=1),                # 162 items: clusters of 2
]


@pytest.mark.parametrize("case", range(len(CONV_CASES)))
def test_conv2d(case):
    conv_case(seed=case, **CONV_CASES[case])


def test_conv2d_rejects_bad_arguments():
    from soccernet_calibration_sportlight_b200 import _lib
    x = torch.zeros(1, 4, 4, 64, dtype=torch.float16, device=dev)
    w = torch.zeros(16, 64, dtype=torch.float16, device=dev)
    y = torch.zeros(1, 4, 4, 64, dtype=torch.float16, device=dev)
    with pytest.raises(_lib.CalError):
        ops.conv2d(x, w, None, y, ksize=5, stride=1, cout_rows=16, relu=False)       # weight K mismatch
    with pytest.raises(_lib.CalError):
        ops.conv2d(x.float(), w, None, y, ksize=1, stride=1, cout_rows=16, relu=False)  # dtype
    y_bad = torch.zeros(1, 3, 4, 64, dtype=torch.float16, device=dev)
    with pytest.raises(_lib.CalError):
        ops.conv2d(x, w, None, y_bad, ksize=1, stride=1, cout_rows=16, relu=False)   # output size


@pytest.mark.parametrize("B,H,W,Cc,srcs,relu,bias", [
    (2, 17, 30, 48, [(17, 30), (9, 15), (5, 8)], True, False),
    (1, 135, 240, 48, [(135, 240), (68, 120), (34, 60), (17, 30)], True, False),
    (1, 68, 120, 96, [(68, 120), (68, 120), (34, 60), (17, 30)], True, False),
    (1, 54, 96, 784, [(27, 48), (14, 24), (7, 12), (4, 6)], False, True),
    (1, 1, 1, 48, [(1, 1), (1, 1)], False, False),
])
def test_fuse_combine(B, H, W, Cc, srcs, relu, bias):
    g = torch.Generator().manual_seed(3)
    cp = packing.pad_to(Cc)
    t16 = [packing.to_nhwc16(torch.randn(B, Cc, h, w, generator=g).to(dev)) for (h, w) in srcs]
    bvec = (torch.randn(cp, generator=g) * 0.3).to(dev) if bias else None
    ref = torch.zeros(B, Cc, H, W, device=dev)
    if bias:
        ref = ref + bvec[:Cc].view(1, -1, 1, 1)
    for t in t16:
        tf = packing.from_nhwc16(t, Cc)
        if tf.shape[-2:] != (H, W):
            tf = F.interpolate(tf, size=(H, W), mode="bilinear", align_corners=True)
        ref = ref + tf
    ref = F.relu(ref) if relu else ref
    y = torch.full((B, H, W, cp), float("nan"), dtype=torch.float16, device=dev)
    ops.fuse_combine(y, t16, bvec, relu)
    got = packing.from_nhwc16(y, Cc)
    assert float((got - ref).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,H,W", [(1, 20, 24), (2, 21, 37), (1, 540, 960)])
def test_stem_conv(B, H, W):
    g = torch.Generator().manual_seed(4)
    x = torch.rand(B, 3, H, W, generator=g).to(dev)
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.3).to(dev)
    b = (torch.randn(64, generator=g) * 0.1).to(dev)
    ref = F.relu(F.conv2d(x, w, b, stride=2, padding=1))
    Ho, Wo = ref.shape[-2:]
    y = torch.full((B, Ho, Wo, 64), float("nan"), dtype=torch.float16, device=dev)
    ops.stem_conv(x, w.reshape(64, 27).contiguous(), b, y)
    got = packing.from_nhwc16(y, 64)
    assert float((got - ref).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,H,W,lows,cout", [
    (2, 34, 48, [(17, 24), (9, 12), (5, 6), (3, 3)], 784),      # keypoint-style: x2, x4, x8, x16 below the head
    (1, 270, 480, [(135, 240), (68, 120), (34, 60), (17, 30)], 784),
    (2, 33, 60, [(17, 30), (9, 15), (5, 8)], 720),              # line-style: three lower branches, ragged tiles
])
def test_head_fused_vs_torch(B, H, W, lows, cout):
    """cal_head_fused = relu(W_full * full + sum_i bilinear(p_i) + bias) against torch
    (F.interpolate align_corners=True, the reference's op, hrnet.py:489-506)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    cpad = (cout + 63) // 64 * 64
    full = torch.randn(B, 64, H, W, generator=g)
    full[:, 48:] = 0                                             # padded channel lanes
    wf = torch.randn(cout, 64, generator=g) / 8
    ps = [torch.randn(B, cout, h, w, generator=g) for h, w in lows]
    bias = torch.randn(cout, generator=g)
    full16, wf16 = full.half(), wf.half()
    ps16 = [t.half() for t in ps]
    ref = F.conv2d(full16.float(), wf16.float()[:, :, None, None]) + bias[None, :, None, None]
    for t in ps16:
        ref = ref + F.interpolate(t.float(), size=(H, W), mode="bilinear", align_corners=True)
    ref = ref.relu()

    def nhwc(t, c):
        out = torch.zeros(t.shape[0], t.shape[2], t.shape[3], c, dtype=torch.float16)
        out[..., :t.shape[1]] = t.permute(0, 2, 3, 1)
        return out.to(dev)
    rows = (cout + 15) // 16 * 16
    w_packed = torch.zeros(rows, 64, dtype=torch.float16)
    w_packed[:cout] = wf16
    bp = torch.zeros(cpad)
    bp[:cout] = bias
    z = torch.full((B, H, W, cpad), float("nan"), dtype=torch.float16, device=dev)
    out = ops.head_fused(nhwc(full16, 64), w_packed.to(dev), [nhwc(t, cpad) for t in ps16], bp.to(dev), z, rows)
    assert out is not None
    got = out[..., :cout].permute(0, 3, 1, 2).float().cpu()
    assert bool(torch.isfinite(out).all())
    assert float(out[..., cout:].abs().max()) == 0.0             # pad lanes stay zero
    err = (got - ref).abs()
    # fp16 interpolation weights (2^-11 relative) + fp16 output rounding
    assert float(err.max()) <= 2e-2 * max(1.0, float(ref.abs().max())) / 4
    assert float(err.mean()) <= 2e-3

    # chained tail: final 1x1 conv + (Log)Softmax on chip (hrnet.py:325-329 / line/hrnet.py:97-101)
    ncls, mode = (58, 1) if cout == 784 else (23, 2)
    w2 = (torch.randn(ncls, cout, generator=g) / 16).half()
    b2 = torch.randn(ncls, generator=g)
    zq = got.half().float()                                        # what the kernel feeds the second GEMM
    logits = F.conv2d(zq, w2.float()[:, :, None, None]) + b2[None, :, None, None]
    ref2 = torch.log_softmax(logits, 1) if mode == 1 else torch.softmax(logits, 1)
    w2p = torch.zeros(64, cpad, dtype=torch.float16)
    w2p[:ncls, :cout] = w2
    b2p = torch.zeros(64)
    b2p[:ncls] = b2
    heat = torch.full((B, ncls, H, W), float("nan"), device=dev)
    out2 = ops.head_fused(nhwc(full16, 64), w_packed.to(dev), [nhwc(t, cpad) for t in ps16], bp.to(dev), None, rows,
                          w2=w2p.to(dev), bias2=b2p.to(dev), heat=heat, mode=mode)
    assert out2 is not None and bool(torch.isfinite(out2).all())
    e2 = (out2.cpu() - ref2).abs()
    assert float(e2.max()) <= (0.05 if mode == 1 else 5e-3), float(e2.max())


BLOCK_CASES = [
    dict(B=1, H=8, W=16, c=48),
    dict(B=2, H=13, W=50, c=48),            # ragged rows and columns, two strips per frame
    dict(B=1, H=135, W=240, c=48),          # the network's own shape
    dict(B=3, H=21, W=61, c=32),
    dict(B=2, H=17, W=30, c=18),            # w18 config: 18 channels in 32 weight rows
    dict(B=40, H=9, W=140, c=48),           # 200 strips: several per CTA, every ring slot sees every chunk position
    dict(B=1, H=4, W=28, c=16),
]


@pytest.mark.parametrize("case", BLOCK_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_basicblock_matches_the_two_launches(case):
    """cal_basicblock (conv - ReLU - conv - add - ReLU with the intermediate in shared memory) against the two
    cal_conv2d launches it replaces (bit for bit with CAL_BB_TAP3=0, where it issues the same MMAs in the same
    order), and against fp32."""
    B, H, W, c = case["B"], case["H"], case["W"], case["c"]
    g = torch.Generator().manual_seed(7 + H + W)
    x = torch.randn(B, c, H, W, generator=g)
    ws, bs, packed = [], [], []
    for _ in range(2):
        w = torch.randn(c, c, 3, 3, generator=g) * (1.0 / (c * 9) ** 0.5)
        b = torch.randn(c, generator=g) * 0.1
        wp, bp, rows = packing.pack_conv(w.double(), b.double())
        packed.append((wp.reshape(rows, -1, 64).permute(1, 0, 2).contiguous().to(dev), bp.to(dev), rows))
        ws.append(w.to(torch.float16).float().to(dev)); bs.append(b.to(dev))
    xh = packing.to_nhwc16(x.to(dev))
    rows = packed[0][2]
    t = torch.full_like(xh, float("nan"))
    ops.conv2d(xh, packed[0][0], packed[0][1], t, ksize=3, stride=1, cout_rows=rows, relu=True, cin=c, w_slices=True)
    y_two = torch.full_like(xh, float("nan"))
    ops.conv2d(t, packed[1][0], packed[1][1], y_two, ksize=3, stride=1, cout_rows=rows, relu=True, res=xh, cin=c, w_slices=True)
    y = torch.full_like(xh, float("nan"))
    ops.basicblock(xh, packed[0][0], packed[0][1], packed[1][0], packed[1][1], y, rows=rows, c=c)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(y).all())
    assert bool((y[..., c:] == 0).all()), "channel padding lanes must be written as zero"
    diff = float((y.float() - y_two.float()).abs().max())
    if os.environ.get("CAL_BB_TAP3", "0") != "1":
        assert torch.equal(y, y_two), f"max |diff| {diff:.3e}"      # filter-row grouping of conv3x3.cu: same MMAs, same order
    else:
        # three taps per MMA: the taps of a filter row are summed after, not inside, the accumulation - fp32 rounding
        # differences that survive the fp16 store as at most an ulp or two
        assert diff <= 2e-3 * max(1.0, float(y_two.float().abs().max())), f"max |diff| {diff:.3e}"
    xr = packing.from_nhwc16(xh, c)
    mid = F.relu(F.conv2d(xr, ws[0], bs[0], padding=1)).to(torch.float16).float()
    ref = F.relu(F.conv2d(mid, ws[1], bs[1], padding=1) + xr)
    err = float((packing.from_nhwc16(y, c) - ref).abs().max())
    assert err <= 4e-3 * max(1.0, float(ref.abs().max())) + 2e-3, f"max err {err:.3e}"


@pytest.mark.parametrize("c,cp", [(48, 64), (96, 128), (18, 64), (40, 64)])
def test_fuse_combine_skips_pad_lanes(c, cp):
    """With the real channel count given, the pad lanes are written as zeros without being gathered: same bits as the
    full-width launch on sources whose pad lanes are zero."""
    g = torch.Generator().manual_seed(c)
    B, H, W = 2, 19, 37
    srcs = []
    for (h, w) in ((H, W), (10, 19), (5, 10)):
        t = torch.zeros(B, h, w, cp, dtype=torch.float16)
        t[..., :c] = torch.randn(B, h, w, c, generator=g).half()
        srcs.append(t.to(dev))
    y0 = torch.full((B, H, W, cp), float("nan"), dtype=torch.float16, device=dev)
    y1 = torch.full_like(y0, float("nan"))
    ops.fuse_combine(y0, srcs, None, relu=True)
    ops.fuse_combine(y1, srcs, None, relu=True, c=c)
    assert bool(torch.isfinite(y1).all()) and bool((y1[..., c:] == 0).all())
    assert torch.equal(y0, y1)
