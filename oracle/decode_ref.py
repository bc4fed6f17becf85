"""ORACLE (test infrastructure, never shipped or timed as product): CPU restatement
of the reference's two heat-map decoders.

* keypoints: ``src/models/hrnet/transforms.py:224-239`` (HRNetPredictionTransform)
* lines:     ``src/models/line/transforms.py:216-280`` (EHMPredictionTransform) and
             ``src/utils/export_line_result.py:51-131`` (get_line_data,
             calculate_slope_intercept), consumer ``src/models/hrnet/prediction.py:643-653``.

Two flavours per decoder:
  *_torch  - literal restatement with the same torch calls the reference makes
             (what the reference computes on CPU).
  *_np     - numpy restatement with every rounding step spelled out; ``exp`` is the
             correctly rounded fp32 exponential (computed in fp64 then rounded), which
             is the definition the CUDA kernels implement.  torch's CPU ``exp`` (SLEEF,
             <=1 ulp) may differ from it in the last bit; that can only move an index
             when two candidates are within a few ulp of each other.  The golden
             generator measures the agreement of both flavours with the real reference.

Pinned by ``tests/golden/make_golden_decode.py`` (reference outputs committed under
tests/golden/decode_*.npz).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

try:  # torch is only needed by the *_torch flavours
    import torch
except Exception:  # pragma: no cover
    torch = None


def exp_f32(x: np.ndarray) -> np.ndarray:
    """Correctly rounded fp32 exp (fp64 evaluation, one final rounding)."""
    return np.exp(np.asarray(x, dtype=np.float32).astype(np.float64)).astype(np.float32)


# ----------------------------------------------------------------------------- keypoints
def keypoint_decode_torch(logp: "torch.Tensor", size=(540, 960)) -> "torch.Tensor":
    """Literal transforms.py:228-239."""
    H, W = size
    _, _, h, w = logp.shape
    p = torch.exp(logp)
    x_prob, x = torch.max(torch.max(p, dim=2)[0], dim=2)
    y_prob, y = torch.max(torch.max(p, dim=3)[0], dim=2)
    conf = torch.min(x_prob, y_prob)
    x = x * W / w
    y = y * H / h
    return torch.stack([x, y, conf], dim=-1)[:, :-1, :]


def keypoint_decode_np(logp: np.ndarray, size=(540, 960)) -> np.ndarray:
    """(B,C,h,w) fp32 log-probs -> (B,C-1,3) fp32 [x, y, conf].

    x: first index of the max over columns of the column-maxima; y: likewise over
    rows; both after exp (ties created by exp saturation resolve to the FIRST
    index, torch CPU semantics); conf = min(x_prob, y_prob)."""
    logp = np.asarray(logp, dtype=np.float32)
    B, C, h, w = logp.shape
    H, W = size
    p = exp_f32(logp)
    col = p.max(axis=2)                       # (B,C,w)
    row = p.max(axis=3)                       # (B,C,h)
    xi = col.argmax(axis=2)                   # numpy argmax = first occurrence
    yi = row.argmax(axis=2)
    xp = np.take_along_axis(col, xi[..., None], 2)[..., 0]
    yp = np.take_along_axis(row, yi[..., None], 2)[..., 0]
    conf = np.minimum(xp, yp)
    # int64 * W -> fp32, then fp32 division by w (torch true-division of int tensors)
    x = (xi * W).astype(np.float32) / np.float32(w)
    y = (yi * H).astype(np.float32) / np.float32(h)
    out = np.stack([x, y, conf], axis=-1).astype(np.float32)
    return out[:, :-1, :]


# ----------------------------------------------------------------------------- lines
def line_decode_torch(heat: "torch.Tensor", sigma: float = 5) -> "torch.Tensor":
    """Literal line/transforms.py:224-280 (mask_heat_points_gauss), CPU."""
    B, C, H, W = heat.shape
    out = -torch.ones(B, C, 2, 3)
    heat = torch.relu(heat)
    xs = torch.arange(0, W, dtype=torch.float)[None, :]
    ys = torch.arange(0, H, dtype=torch.float)[:, None]
    for b in range(B):
        for c in range(C):
            hm = heat[b, c]
            v1, i1 = torch.max(hm.view(-1), dim=0)
            c1 = torch.tensor([i1 % W, i1 // W], dtype=torch.float)
            out[b, c, 0, :2] = c1
            out[b, c, 0, 2] = v1
            mask = torch.exp(-((xs - c1[0]) ** 2 + (ys - c1[1]) ** 2) / (2.0 * sigma ** 2))
            hm2 = hm * (1 - mask)
            v2, i2 = torch.max(hm2.view(-1), dim=0)
            out[b, c, 1, :2] = torch.tensor([i2 % W, i2 // W], dtype=torch.float)
            out[b, c, 1, 2] = v2
    return out


def line_decode_np(heat: np.ndarray, sigma: float = 5) -> np.ndarray:
    """(B,C,H,W) fp32 probabilities -> (B,C,2,3) fp32: two peaks [x, y, value] per
    channel; the second is the flat first-argmax of relu(h) * (1 - G) with
    G = exp(-(dx^2+dy^2) / (2 sigma^2)) centred on the first, all in fp32."""
    heat = np.maximum(np.asarray(heat, dtype=np.float32), np.float32(0))
    B, C, H, W = heat.shape
    out = np.empty((B, C, 2, 3), dtype=np.float32)
    xs = np.arange(W, dtype=np.float32)[None, :]
    ys = np.arange(H, dtype=np.float32)[:, None]
    denom = np.float32(2.0 * sigma ** 2)
    one = np.float32(1)
    for b in range(B):
        for c in range(C):
            hm = heat[b, c]
            i1 = int(hm.reshape(-1).argmax())
            x1, y1 = np.float32(i1 % W), np.float32(i1 // W)
            out[b, c, 0] = (x1, y1, hm.reshape(-1)[i1])
            d2 = (xs - x1) ** 2 + (ys - y1) ** 2              # exact small integers in fp32
            mask = exp_f32((-d2) / denom)
            hm2 = hm * (one - mask)
            i2 = int(hm2.reshape(-1).argmax())
            out[b, c, 1] = (np.float32(i2 % W), np.float32(i2 // W), hm2.reshape(-1)[i2])
    return out


def line_transform_np(heat: np.ndarray, scale: float = 8, sigma: float = 6) -> np.ndarray:
    """EHMPredictionTransform.__call__ (line/transforms.py:216-222): decode then scale x, y."""
    pred = line_decode_np(heat, sigma)
    pred[..., 0] *= np.float32(scale)
    pred[..., 1] *= np.float32(scale)
    return pred


def calculate_slope_intercept(p1, p2, delta: float = 0.00001):
    """export_line_result.py:51-82."""
    if tuple(p1) == tuple(p2):
        return None, None
    x1, y1 = p1
    x2, y2 = p2
    slope = (y2 - y1) / (x2 - x1 + delta)
    return slope, y1 - slope * x1


def get_line_data(heat_loc: np.ndarray, line_cls: Dict[int, str], scale=4, prob_thre: float = 0.2):
    """export_line_result.py:85-131 (frame 0 of the batch only, as there).  The reference pins
    numpy==1.24.2 (requirements.txt:2), where a float32 scalar times the Python int ``scale``
    promotes to float64 (legacy scalar promotion; numpy 2 would stay in float32): the coordinates,
    and with them slope, intercept and the intersections downstream, are float64."""
    heat_loc = np.asarray(heat_loc)
    _, ks, nh, _ = heat_loc.shape
    lines: Dict[str, Tuple[float, float]] = {}
    points: Dict[str, List[Tuple[float, float, float]]] = {}
    for k in range(ks):
        valid = []
        for n in range(nh):
            x, y, p = heat_loc[0, k, n]
            if p >= prob_thre:
                valid.append((np.float64(x) * scale, np.float64(y) * scale, p))
        points[line_cls[k]] = valid
        if len(valid) >= 2:
            lines[line_cls[k]] = calculate_slope_intercept(valid[0][:2], valid[1][:2])
    return lines, points


def line_eq_intersection(l1, l2) -> Optional[Tuple[float, float]]:
    """prediction.py:643-653."""
    k1, b1 = l1
    k2, b2 = l2
    if abs(k1 - k2) > 1e-4:
        x = (b2 - b1) / (k1 - k2)
        return x, k1 * x + b1
    return None


def lines_to_keypoints(lines: Dict[str, Tuple[float, float]], line_intersections) -> Dict[int, Tuple[float, float]]:
    """CameraCreator.__init__ line handling (prediction.py:110-124): per line pair of
    LINE_INTERSECTIONS present -> keypoint id -> (x, y)."""
    pts = {}
    for idx, (a, b) in line_intersections.items():
        if a in lines and b in lines:
            la, lb = lines[a], lines[b]
            if la[0] is None or lb[0] is None:
                continue
            p = line_eq_intersection(la, lb)
            if p is not None:
                pts[idx] = p
    return pts
