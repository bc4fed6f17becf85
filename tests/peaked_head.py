"""Shared by the end-to-end parity tests: turns the seeded oracle network into one whose heat maps have
confident, blob-shaped peaks on a given frame - with plain random-init weights every confidence is
~1/58 and the argmax says nothing about what half-precision operands do to the keypoints the path
ships.  Two steps, both on the fp32 CPU oracle (the GPU network then loads the very
same state dict):

1. ``calibrate_bn``: the BatchNorm running statistics are set to the statistics of this frame (one
   training-mode forward with momentum 1).  Random BN statistics (hrnet_ref.randomize_bn_) leave the
   features ~99 % input-independent; calibrated, every layer's output is zero-mean / unit-variance over
   the frame, as in a trained network, and the features are a (random) function of the local image.
2. ``install_templates``: the last 1x1 conv (hrnet.py:325-329) is rewritten so that keypoint c's logit
   is contrast * (<z(p) - zbar, z(p_c) - zbar> / |z(p_c) - zbar|^2 - 1) + peak_logit: a smooth blob
   around the wanted heat-map position p_c (neighbouring pixels ~0.05 * contrast lower), far below the
   constant background logit elsewhere.  All 306 convolutions before it are untouched, so the hidden
   features - and their half-precision error on the GPU - are those of the seeded network.
"""
from __future__ import annotations

import numpy as np
import torch


@torch.no_grad()
def calibrate_bn(oracle, x: torch.Tensor) -> None:
    mods = [m for m in oracle.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    old = [m.momentum for m in mods]
    for m in mods:
        m.momentum = 1.0
    oracle.train()
    oracle(x)
    oracle.eval()
    for m, mom in zip(mods, old):
        m.momentum = mom


@torch.no_grad()
def install_templates(oracle, x: torch.Tensor, positions, peak_logit: float = 10.0, contrast: float = 8.0,
                      bg_logit: float = 5.0):
    """oracle: hrnet_ref.HRNetHeatmapRef (keypoints); x: (1,3,H,W) frame; positions: {channel: (row, col)} in
    heat-map pixels.  Rewrites last_layer[3] in place; returns the hidden tensor's shape."""
    net = oracle.model
    _, cat = net(x)
    z = net.last_layer[2](net.last_layer[1](net.last_layer[0](cat)))[0]       # (C, h, w) after ReLU
    C, h, w = z.shape
    zbar = z.reshape(C, -1).mean(1)
    conv = net.last_layer[3]
    W = torch.zeros_like(conv.weight)                                         # (58, C, 1, 1)
    b = torch.full_like(conv.bias, -30.0)
    b[-1] = bg_logit
    for c, (r, q) in positions.items():
        d = z[:, r, q] - zbar
        s = contrast / float(d @ d)
        W[c, :, 0, 0] = s * d
        b[c] = -s * float(d @ zbar) - contrast + peak_logit
    conv.weight.copy_(W)
    conv.bias.copy_(b)
    return (C, h, w)


def camera_positions(seed: int, h: int = 270, w: int = 480, min_sep: int = 12):
    """Heat-map positions of the pitch keypoints seen by a plausible broadcast camera (tests/camera_inputs),
    thinned so that no two are closer than `min_sep` heat-map pixels."""
    from tests import camera_inputs as CI
    rng = np.random.default_rng(seed)
    for _ in range(200):
        R, pos, f = CI.random_camera(rng)
        uv, ok = CI.project(R, pos, f)
        out = {}
        for c in range(57):
            if not ok[c]:
                continue
            r, q = int(round(uv[c, 1] / 2)), int(round(uv[c, 0] / 2))
            if not (4 <= r < h - 4 and 4 <= q < w - 4):
                continue
            if all(max(abs(r - a), abs(q - bq)) >= min_sep for a, bq in out.values()):
                out[c] = (r, q)
        if len(out) >= 14:
            return out, (R, pos, f)
    raise RuntimeError("no camera with enough visible keypoints")
