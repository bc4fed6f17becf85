"""Host-side weight preparation: fold eval-mode BatchNorm into the preceding conv
(hrnet.py:35-38, 67-74, 149, 191, 203, 211, 262-266, 321: y = gamma*(conv-mu)/sqrt(var+eps)+beta)
and repack to the kernel layouts (fp16, K-major [Cout][tap][Cin_pad]; fp32 bias)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

BN_EPS = 1e-5  # torch.nn.BatchNorm2d / SyncBatchNorm default, never overridden by the reference


def pad_to(c: int, m: int = 64) -> int:
    return (c + m - 1) // m * m


def fold_bn(w: torch.Tensor, b: Optional[torch.Tensor], bn: Optional[dict]) -> Tuple[torch.Tensor, torch.Tensor]:
    """w (Co,Ci,k,k), optional conv bias (Co), bn = dict(weight,bias,running_mean,running_var)
    or None.  Folding is done in fp64, returned as fp64 (rounded once when packed)."""
    w = w.detach().double().cpu()
    co = w.shape[0]
    b = torch.zeros(co, dtype=torch.float64) if b is None else b.detach().double().cpu()
    if bn is None:
        return w, b
    g = bn["weight"].detach().double().cpu()
    beta = bn["bias"].detach().double().cpu()
    mu = bn["running_mean"].detach().double().cpu()
    var = bn["running_var"].detach().double().cpu()
    s = g / torch.sqrt(var + BN_EPS)
    return w * s.view(-1, 1, 1, 1), (b - mu) * s + beta


def pack_conv(w: torch.Tensor, b: torch.Tensor, cin_pad: Optional[int] = None, cout_pad: Optional[int] = None,
              cin_offset: int = 0, cin_total_pad: Optional[int] = None):
    """-> (w_packed fp16 (rows, k*k*cin_pad), bias fp32 (cout_pad), rows).
    rows = Cout rounded up to 16 (MMA N granularity); pad rows/columns are zero."""
    co, ci, kh, kw = w.shape
    cin_pad = pad_to(ci) if cin_pad is None else cin_pad
    cout_pad = pad_to(co) if cout_pad is None else cout_pad
    rows = pad_to(co, 16)
    wp = torch.zeros(rows, kh * kw, cin_pad, dtype=torch.float64)
    wp[:co, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
    bp = torch.zeros(cout_pad, dtype=torch.float64)
    bp[:co] = b
    return wp.reshape(rows, kh * kw * cin_pad).to(torch.float16).contiguous(), bp.to(torch.float32).contiguous(), rows


def to_nhwc16(x: torch.Tensor, c_pad: Optional[int] = None) -> torch.Tensor:
    """fp32 NCHW -> fp16 NHWC with zero channel padding."""
    B, Cc, H, W = x.shape
    c_pad = pad_to(Cc) if c_pad is None else c_pad
    y = torch.zeros(B, H, W, c_pad, dtype=torch.float16, device=x.device)
    y[..., :Cc] = x.permute(0, 2, 3, 1).to(torch.float16)
    return y


def from_nhwc16(y: torch.Tensor, c: int) -> torch.Tensor:
    """fp16 NHWC (padded) -> fp32 NCHW with the first c channels."""
    return y[..., :c].permute(0, 3, 1, 2).float().contiguous()
