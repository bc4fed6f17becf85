"""Shared parity bookkeeping for the camera solve: compares camera records produced by a solver
under test (the CUDA kernel on the GPU box, or the same source compiled for the host in the CPU
suite) with the stored outputs of the UNMODIFIED reference (tests/golden/camera_cases.npz,
made by tests/golden/make_golden_camera.py).

Parity classes (DESIGN.md "camera solve: parity classes"):
  exact     - the reference's outcome is a deterministic function of the input and goes only
              through routines whose result is a well-defined minimiser (calibrateCamera,
              solvePnPRefineLM, findHomography): camera parameters must agree within 1e-4
              relative, None decisions must agree;
  ransac    - the outcome consulted a camera that went through a SUCCESSFUL solvePnPRansac with
              more than 5 points: deterministic in the reference, but seeded by EPnP on random
              5-point samples of OpenCV's internal RNG - on the near-coplanar pitch points often the
              flipped planar pose, which refine_camera keeps; where OpenCV lands on the
              least-squares pose the results agree; reported, not asserted;
  minimal   - the outcome went through solvePnPRansac with exactly 5 (or 4) points (OpenCV returns
              its EPnP (P3P) minimal solver's answer, typically the flipped planar pose):
              deterministic in the reference but not restated; reported, not asserted;
  unpinned  - the outcome went through a FAILED solvePnPRansac: the reference consumes
              uninitialised memory and does not reproduce itself; reported, not asserted;
  ill-posed - exact-class frames whose least-squares problem has no well-defined minimiser:
              (a) the reference's own camera fails its feasibility gate (good_camera,
              prediction.py:469-475: focal length running away along a flat valley that OpenCV
              leaves after 30 iterations, or a local minimum with the camera under the pitch -
              only opencv_calibration[_multiplane] return such cameras at all), (b) a goal-plane view made of
              collinear goal-line points only (no crossbar point): the view's homography is
              undetermined and the minimum reached depends on OpenCV's internal start.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict

import numpy as np

from soccernet_calibration_sportlight_b200 import _lib, pitch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "camera_cases.npz")
SETS = ("noisy", "clean", "sparse", "wide")
ALGOS = {"opencv_calibration": 0.5, "opencv_calibration_multiplane": 0.5, "original_voter": 0.5, "voter": 0.35,
         "iterative_voter": 0.5}
ALGO_ID = {"opencv_calibration": 0, "opencv_calibration_multiplane": 1, "original_voter": 2, "voter": 3,
           "iterative_voter": 4}
TOL = 1e-4          # relative, on position (metres, floor 1), rotation entries (floor 1) and focal lengths


def make_params(algo: str, thr: float) -> "_lib.SolveParams":
    P = _lib.SolveParams()
    for i, v in enumerate(pitch.keypoint_world_table().reshape(-1)):
        P.pitch_xyz[i] = float(v)
    P.algorithm, P.img_w, P.img_h, P.conf_thresh = ALGO_ID[algo], 960, 540, thr
    for i, v in enumerate((0.5, 0.35, 0.2)):
        P.conf_threshs[i] = v
    P.n_conf_threshs = 3
    P.min_points, P.min_points_per_plane, P.min_points_for_refinement, P.reliable_thresh = 5, 6, 6, 57
    P.min_focal_length, P.max_rmse, P.max_rmse_rel = 10.0, 55.0, 5.0
    return P


def records_to_array(rec) -> np.ndarray:
    """ctypes CameraRecord array -> (B,16): position, rotation, fx, fy, valid, branch."""
    out = np.zeros((len(rec), 16))
    for i, r in enumerate(rec):
        if r.valid:
            out[i, 0:3], out[i, 3:12], out[i, 12], out[i, 13], out[i, 14] = list(r.position), list(r.rotation), r.fx, r.fy, 1.0
        out[i, 15] = r.branch
    return out


def degenerate_goal_view(pred: np.ndarray, thr: float) -> bool:
    """A goal-plane view with >= 6 selected points of which fewer than two are crossbar ends: at
    least five points on the goal line (collinear in the view's plane coordinates) and at most
    one off it - no four points in general position, so the view's homography is undetermined."""
    sel = {i for i in range(57) if pred[i, 2] > thr}
    for ids, tops in ((pitch.POINT_SETS["goal_left"], (0, 1)), (pitch.POINT_SETS["goal_right"], (24, 25))):
        got = [i for i in ids if i in sel]
        if len(got) >= 6 and sum(t in got for t in tops) < 2:
            return True
    return False


def feasible_record(rec: np.ndarray) -> bool:
    """good_camera (prediction.py:469-475) on a camera record."""
    return bool(10 <= rec[12] <= 20000 and -250 < rec[0] < 250 and -250 < rec[1] < 250 and -100 < rec[2] < 0)


def classify(z, sname: str, algo: str, i: int) -> str:
    if not bool(z[f"{sname}__{algo}__pinned"][i]):
        return "unpinned"
    if bool(z[f"{sname}__{algo}__minimal"][i]):
        return "minimal"
    if bool(z[f"{sname}__{algo}__ransac"][i]):
        return "ransac"
    ref = z[f"{sname}__{algo}__records"][i]
    if ref[14] == 1 and not feasible_record(ref):
        return "ill-posed"
    thrs = (0.5, 0.35, 0.2) if algo == "iterative_voter" else (ALGOS[algo],)
    if any(degenerate_goal_view(z[f"{sname}__preds"][i], t) for t in thrs):
        return "ill-posed"
    return "exact"


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    scale = np.maximum(np.abs(ref[:14]), 1.0)
    return float(np.max(np.abs(got[:14] - ref[:14]) / scale))


def compare(solver: Callable[[np.ndarray, str, float], np.ndarray], algos=None, sets=SETS) -> Dict[str, dict]:
    """Runs ``solver(preds, algo, thr) -> (B,16)`` over the golden cases; returns per class
    {'n', 'decision_agree', 'param_agree', 'max_err', 'failures': [...]}."""
    z = np.load(GOLDEN)
    stats: Dict[str, dict] = {}
    for sname in sets:
        preds = np.ascontiguousarray(z[f"{sname}__preds"])
        for algo in (algos or ALGOS):
            got = solver(preds, algo, ALGOS[algo])
            ref = z[f"{sname}__{algo}__records"]
            for i in range(preds.shape[0]):
                cls = classify(z, sname, algo, i)
                s = stats.setdefault(cls, dict(n=0, decision_agree=0, param_agree=0, max_err=0.0, failures=[]))
                s["n"] += 1
                if (got[i, 14] == 1) != (ref[i, 14] == 1):
                    s["failures"].append((sname, algo, i, "decision", float(got[i, 14]), float(ref[i, 14])))
                    continue
                s["decision_agree"] += 1
                if got[i, 14] == 1:
                    e = rel_err(got[i], ref[i])
                    if not np.isfinite(e) or e > TOL:
                        s["failures"].append((sname, algo, i, "params", e, str(z[f"{sname}__{algo}__branch"][i])))
                        continue
                    s["max_err"] = max(s["max_err"], e)
                s["param_agree"] += 1
    return stats


def host_solver():
    """The CUDA solver's source compiled for the host (tests/host_solver): test infrastructure to
    check the numerics on CPU-only machines.  Returns solver(preds, algo, thr) -> (B,16)."""
    import subprocess
    src = os.path.join(ROOT, "tests", "host_solver", "host_solver.cpp")
    so = os.path.join(ROOT, "tests", "host_solver", "libhost_solver.so")
    deps = [src] + [os.path.join(ROOT, "soccernet_calibration_sportlight_b200", "csrc", f)
                    for f in ("solve_core.cuh", "solve_cascade.cuh")] + [os.path.join(ROOT, "include", "calib_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    L = C.CDLL(so)

    def solve(preds, algo, thr, line_pts=None):
        preds = np.ascontiguousarray(preds, dtype=np.float32)
        rec = (_lib.CameraRecord * preds.shape[0])()
        P = make_params(algo, thr)
        lp = None if line_pts is None else np.ascontiguousarray(line_pts, dtype=np.float64).ctypes.data_as(C.c_void_p)
        L.host_camera_solve(preds.ctypes.data_as(C.c_void_p), lp, C.byref(P), preds.shape[0], rec)
        return records_to_array(rec)
    return solve
