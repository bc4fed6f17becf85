"""Deterministic synthetic inputs shared by the golden-vector generators (build
container) and the tests (CPU and GPU box).  Everything here is exact integer /
dyadic-rational arithmetic, so the arrays are bit-identical on any machine and
need not be committed; only the reference's OUTPUTS on them are stored under
tests/golden/."""
from __future__ import annotations

import numpy as np


def _hash32(a: np.ndarray, seed: int) -> np.ndarray:
    """xorshift-multiply integer hash on uint64 lanes, returns uint32."""
    with np.errstate(over="ignore"):      # wrap-around mod 2^64 is the point
        x = a.astype(np.uint64) + np.uint64((0x9E3779B97F4A7C15 * (seed + 1)) & 0xFFFFFFFFFFFFFFFF)
        x ^= x >> np.uint64(30)
        x = x * np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x = x * np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def hashed_logp(seed: int, shape, bits: int = 18, frac: int = 14) -> np.ndarray:
    """'Noise' log-probabilities: -(hash mod 2^bits) * 2^-frac, exactly representable
    in fp32 (range (-16, 0] for the defaults).  Many near-ties and exp-collisions."""
    n = int(np.prod(shape))
    h = _hash32(np.arange(n, dtype=np.uint64), seed) & np.uint32((1 << bits) - 1)
    return (-(h.astype(np.float64)) * 2.0 ** -frac).astype(np.float32).reshape(shape)


def gaussian_logp(seed: int, B: int, C: int, h: int, w: int, inv2s2_log2: int = 4,
                  floor: float = -40.0, invisible_every: int = 7, halfpix_every: int = 5) -> np.ndarray:
    """Log of a unit-peak Gaussian per channel: logp = max(floor, -(dx^2+dy^2) * 2^-k).
    Centres are hashed; every ``halfpix_every``-th channel is centred on a half
    pixel (4-way exact tie), every ``invisible_every``-th channel is flat ``floor``
    (an 'invisible' keypoint: all-equal plane -> index 0).  Values near the peak
    saturate exp() to 1.0f, producing plateaus; all arithmetic is exact in fp32."""
    out = np.empty((B, C, h, w), dtype=np.float32)
    ys = np.arange(h, dtype=np.float64)[:, None]
    xs = np.arange(w, dtype=np.float64)[None, :]
    hh = _hash32(np.arange(B * C * 2, dtype=np.uint64), seed).reshape(B, C, 2)
    for b in range(B):
        for c in range(C):
            if invisible_every and (b * C + c) % invisible_every == invisible_every - 1:
                out[b, c] = floor
                continue
            cx = float(hh[b, c, 0] % np.uint32(w))
            cy = float(hh[b, c, 1] % np.uint32(h))
            if halfpix_every and (b * C + c) % halfpix_every == 0:
                cx = min(cx, w - 2) + 0.5
                cy = min(cy, h - 2) + 0.5
            d2 = (xs - cx) ** 2 + (ys - cy) ** 2          # multiples of 0.25: exact
            out[b, c] = np.maximum(floor, -d2 * 2.0 ** -inv2s2_log2).astype(np.float32)
    return out


def two_peak_heat(seed: int, B: int, C: int, h: int, w: int) -> np.ndarray:
    """Line-model style probabilities in [0,1): a hashed background (multiples of
    2^-12 below 1/16) plus two dyadic 'tent' peaks per channel; a few channels are
    empty / negative to exercise the relu and the all-zero plane."""
    n = B * C * h * w
    bg = (_hash32(np.arange(n, dtype=np.uint64), seed) & np.uint32(0xFF)).astype(np.float64) * 2.0 ** -12
    out = bg.reshape(B, C, h, w)
    ys = np.arange(h, dtype=np.float64)[:, None]
    xs = np.arange(w, dtype=np.float64)[None, :]
    hh = _hash32(np.arange(B * C * 6, dtype=np.uint64), seed + 17).reshape(B, C, 6)
    for b in range(B):
        for c in range(C):
            k = (b * C + c) % 9
            if k == 8:
                out[b, c] = -out[b, c]                     # all non-positive -> relu -> zeros
                continue
            for j in range(2):
                cx = float(hh[b, c, 3 * j] % np.uint32(w))
                cy = float(hh[b, c, 3 * j + 1] % np.uint32(h))
                amp = 0.5 + float(hh[b, c, 3 * j + 2] % np.uint32(64)) / 256.0
                if k == 7 and j == 1:
                    amp = out[b, c].max()                 # exact tie between the two peaks
                tent = np.maximum(0.0, amp - (np.abs(xs - cx) + np.abs(ys - cy)) * 2.0 ** -4)
                out[b, c] = np.maximum(out[b, c], tent)
    return out.astype(np.float32)


def frames_u8(seed: int, B: int, H: int = 540, W: int = 960) -> np.ndarray:
    """(B,H,W,3) uint8 BGR frames: smooth integer gradients + hashed texture."""
    n = B * H * W * 3
    tex = (_hash32(np.arange(n, dtype=np.uint64), seed) & np.uint32(0x3F)).reshape(B, H, W, 3)
    yy = (np.arange(H, dtype=np.uint32)[None, :, None, None] * np.uint32(3)) // np.uint32(11)
    xx = (np.arange(W, dtype=np.uint32)[None, None, :, None] * np.uint32(5)) // np.uint32(29)
    cc = np.arange(3, dtype=np.uint32)[None, None, None, :] * np.uint32(23)
    return ((tex + yy + xx + cc) % np.uint32(256)).astype(np.uint8)


def frames_to_tensor(frames: np.ndarray) -> np.ndarray:
    """cv2.imread + T.ToTensor (transforms.py:59-68, make_submit.py:66): HWC uint8 ->
    CHW fp32 in [0,1] (division by 255 in fp32)."""
    return (np.transpose(frames, (0, 3, 1, 2)).astype(np.float32) / np.float32(255.0))
