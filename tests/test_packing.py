"""Host-side weight preparation (packing.py): BN folding equals eval-mode BatchNorm, and
the packed layout is [Cout_rows][tap][Cin_pad] with zero padding."""
import torch
import torch.nn as nn

from soccernet_calibration_sportlight_b200 import packing


def test_fold_bn_equals_eval_batchnorm():
    g = torch.Generator().manual_seed(0)
    conv = nn.Conv2d(5, 7, 3, padding=1, bias=True)
    bn = nn.BatchNorm2d(7)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(7, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(7, generator=g))
        bn.running_mean.copy_(torch.randn(7, generator=g))
        bn.running_var.copy_(torch.rand(7, generator=g) + 0.5)
    bn.eval()
    x = torch.randn(2, 5, 6, 8, generator=g)
    ref = bn(conv(x)).double()
    w, b = packing.fold_bn(conv.weight, conv.bias, dict(weight=bn.weight, bias=bn.bias,
                                                         running_mean=bn.running_mean,
                                                         running_var=bn.running_var))
    got = torch.nn.functional.conv2d(x.double(), w, b, padding=1)
    assert float((got - ref).abs().max()) < 1e-5


def test_pack_layout_and_padding():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float64).reshape(2, 3, 3, 3) / 64
    b = torch.tensor([1.0, 2.0], dtype=torch.float64)
    wp, bp, rows = packing.pack_conv(w, b)
    assert rows == 16 and wp.shape == (16, 9 * 64) and bp.shape == (64,)
    wv = wp.reshape(16, 9, 64)
    for co in range(2):
        for tap in range(9):
            for ci in range(3):
                assert float(wv[co, tap, ci]) == float(w[co, ci, tap // 3, tap % 3])
    assert float(wv[2:].abs().max()) == 0 and float(wv[:, :, 3:].abs().max()) == 0
    assert bp[:2].tolist() == [1.0, 2.0] and float(bp[2:].abs().max()) == 0


def test_nhwc_roundtrip():
    x = torch.randn(2, 5, 4, 6)
    y = packing.to_nhwc16(x)
    assert y.shape == (2, 4, 6, 64) and y.dtype == torch.float16
    assert torch.equal(packing.from_nhwc16(y, 5), x.half().float())
