"""Synthetic keypoint predictions for the camera solve: the 57 pitch
keypoints seen by plausible broadcast cameras, quantised to the even-pixel grid the keypoint
decode emits, with confidences, dropped points, pixel noise and gross outliers.

Used by the golden-vector generator (build container), the parity tests and bench.py.  The
golden file stores the generated arrays themselves, so tests never depend on regenerating
bit-identical random streams on another machine."""
from __future__ import annotations

import numpy as np

from soccernet_calibration_sportlight_b200 import pitch

WORLD = pitch.keypoint_world_table()
W_IMG, H_IMG = 960, 540


def rotation_from_pan_tilt_roll(pan, tilt, roll):
    """world->camera rotation of the SoccerNet convention (baseline/camera.py:7-28, 204-205)."""
    rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, np.cos(tilt), -np.sin(tilt)], [0.0, np.sin(tilt), np.cos(tilt)]])
    return (rz(pan) @ rx @ rz(roll)).T


def random_camera(rng):
    pan = np.deg2rad(rng.uniform(-35, 35))
    tilt = np.deg2rad(rng.uniform(70, 85))
    roll = np.deg2rad(rng.uniform(-1, 1))
    pos = np.array([rng.uniform(-10, 10), rng.uniform(50, 80), rng.uniform(-25, -8)])
    f = rng.uniform(900, 4000)
    return rotation_from_pan_tilt_roll(pan, tilt, roll), pos, f


def project(R, pos, f, pts=WORLD):
    pc = (R @ (pts - pos).T).T
    z = pc[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        uv = pc[:, :2] / z[:, None] * f + np.array([W_IMG / 2.0, H_IMG / 2.0])
    ok = (z > 1e-3) & (uv[:, 0] >= 0) & (uv[:, 0] < W_IMG) & (uv[:, 1] >= 0) & (uv[:, 1] < H_IMG)
    return uv, ok


def synthetic_predictions(n, seed=0, noise_px=1.5, drop=0.15, outlier=0.03, conf_lo=0.15, min_visible=0,
                          wide=False):
    """(n,57,3) float32 [x, y, conf].  Visible keypoints get conf U(conf_lo,1) (so the 0.5 /
    0.35 / 0.2 thresholds of make_submit.py all bite), invisible or dropped ones (0,0,0)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 57, 3), dtype=np.float32)
    k = 0
    while k < n:
        R, pos, f = random_camera(rng)
        if wide:
            f = rng.uniform(600, 1500)
        uv, ok = project(R, pos, f)
        if ok.sum() < min_visible:
            continue
        uv = uv + rng.normal(0.0, noise_px, uv.shape)
        bad = rng.uniform(size=57) < outlier
        uv[bad] += rng.uniform(-80, 80, (int(bad.sum()), 2))
        q = 2.0 * np.floor(uv / 2.0 + 0.5)                      # even-pixel grid of the decode
        ok &= (q[:, 0] >= 0) & (q[:, 0] <= W_IMG - 2) & (q[:, 1] >= 0) & (q[:, 1] <= H_IMG - 2)
        ok &= rng.uniform(size=57) >= drop
        conf = rng.uniform(conf_lo, 1.0, 57)
        out[k, ok, 0] = q[ok, 0]
        out[k, ok, 1] = q[ok, 1]
        out[k, ok, 2] = conf[ok]
        k += 1
    return out


def clean_predictions(n, seed=0, conf_lo=0.55):
    """Noise-free (only even-pixel quantisation), no drops, all visible points confident."""
    return synthetic_predictions(n, seed=seed, noise_px=0.0, drop=0.0, outlier=0.0, conf_lo=conf_lo)
