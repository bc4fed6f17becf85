"""Prediction transforms with the reference's names and call signatures, running the CUDA
decoders (csrc/decode.cu) through the C ABI.

* ``HRNetPredictionTransform`` - src/models/hrnet/transforms.py:224-239
* ``EHMPredictionTransform``   - src/models/line/transforms.py:193-280
* ``get_line_data`` / ``calculate_slope_intercept`` - src/utils/export_line_result.py:51-131
  (host-side dictionary building over the 23x2 decoded peaks; no heat-map arithmetic)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .pitch import LINE_CLS


class HRNetPredictionTransform:
    """``size`` = (H, W) of the image the keypoints are expressed in (transforms.py:225)."""

    def __init__(self, size=(540, 960)):
        self.size = (int(size[0]), int(size[1]))

    def __call__(self, preds: torch.Tensor) -> torch.Tensor:
        """(B, C, h, w) fp32 log-probabilities on the GPU -> (B, C-1, 3) [x, y, conf]."""
        return ops.kp_decode(preds.contiguous(), self.size)


class EHMPredictionTransform:
    def __init__(self, scale=8, sigma=6):
        self.scale = scale
        self.sigma = sigma
        self.distance_threshold = 2 * self.sigma

    def __call__(self, preds: torch.Tensor) -> torch.Tensor:
        return ops.line_decode(preds.contiguous(), float(self.sigma), float(self.scale))

    @staticmethod
    def mask_heat_points_gauss(tensor: torch.Tensor, sigma: float = 5) -> torch.Tensor:
        """(B, C, h, w) fp32 probabilities -> (B, C, 2, 3) two peaks [x, y, value]."""
        return ops.line_decode(tensor.contiguous(), float(sigma), 1.0)


def calculate_slope_intercept(p1, p2, delta: float = 0.00001):
    """export_line_result.py:51-82."""
    if tuple(p1) == tuple(p2):
        return None, None
    x1, y1 = p1
    x2, y2 = p2
    slope = (y2 - y1) / (x2 - x1 + delta)
    return slope, y1 - slope * x1


def get_line_data(heat_loc, scale=4, prob_thre: float = 0.2, frame: int = 0):
    """export_line_result.py:85-131 for one frame of the decoded (B,23,2,3) peaks:
    -> (lines {class: (slope, intercept)}, points {class: [(x, y, p), ...]}).  Coordinates, slopes and
    intercepts are float64, as under the reference's pinned numpy 1.24 (float32 scalar x Python int)."""
    if isinstance(heat_loc, torch.Tensor):
        heat_loc = heat_loc.detach().cpu().numpy()
    heat_loc = np.asarray(heat_loc)
    lines: Dict[str, Tuple[float, float]] = {}
    points: Dict[str, List[Tuple[float, float, float]]] = {}
    for k in range(heat_loc.shape[1]):
        valid = []
        for n in range(heat_loc.shape[2]):
            x, y, p = heat_loc[frame, k, n]
            if p >= prob_thre:
                valid.append((np.float64(x) * scale, np.float64(y) * scale, p))
        points[LINE_CLS[k]] = valid
        if len(valid) >= 2:
            lines[LINE_CLS[k]] = calculate_slope_intercept(valid[0][:2], valid[1][:2])
    return lines, points
