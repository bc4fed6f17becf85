"""Shared parity bookkeeping for the camera solve: compares camera records produced by a solver
under test (the CUDA kernel on the GPU box, or the same source compiled for the host in the CPU
suite) with the stored outputs of the UNMODIFIED reference (tests/golden/camera_cases.npz,
made by tests/golden/make_golden_camera.py).

Parity classes (DESIGN.md "camera solve: parity classes").  ASSERTED (camera parameters within 1e-4
relative, None decisions identical):
  exact     - the outcome went only through calibrateCamera / solvePnPRefineLM / findHomography;
  minimal   - ... through solvePnPRansac with exactly 5 (or 4) points: OpenCV returns its EPnP (P3P)
              minimal solver's answer, typically the mirrored planar pose;
  ransac    - ... through a successful solvePnPRansac on more than 5 points (seeded 5-point samples,
              EPnP on each, refit of the consensus set from the winning sample's pose).
              All three are restated in csrc/solve_pnp_cv.cuh.
REPORTED, not asserted - in each the reference's own output is not a function the inputs determine:
  unpinned    - the outcome went through a FAILED solvePnPRansac: the reference consumes
                uninitialised memory and does not reproduce itself from run to run;
  unconverged - some refine_camera of the frame returned a point that still moves under a tighter
                stopping rule (OpenCV's LM stops on a step < 1e-5): "where the solver stopped";
  ill-posed   - the least-squares problem of calibrateCamera has no well-defined minimiser:
                (a) the reference's own camera fails its feasibility gate (good_camera,
                prediction.py:469-475: focal length running away along a flat valley that OpenCV
                leaves after 30 iterations, or a local minimum with the camera under the pitch),
                for the voters judged on the calibrateCamera stage they share with
                opencv_calibration_multiplane; (b) a goal-plane view made of collinear goal-line
                points only (no crossbar point): the view's homography is undetermined and the
                minimum reached depends on OpenCV's internal start.
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


ASSERTED = ("exact", "minimal", "ransac")


def classify(z, sname: str, algo: str, i: int) -> str:
    if not bool(z[f"{sname}__{algo}__pinned"][i]):
        return "unpinned"
    if bool(z[f"{sname}__{algo}__unconverged"][i]):
        return "unconverged"
    ref = z[f"{sname}__{algo}__records"][i]
    if ref[14] == 1 and not feasible_record(ref):
        return "ill-posed"
    if algo in ("original_voter", "iterative_voter"):
        # their first stage is the calibrateCamera call of opencv_calibration_multiplane on the same
        # points (prediction.py:374-408 vs 194-225); when that camera is infeasible the voter silently
        # falls through to its homography camera - or not, depending on where OpenCV's LM stopped
        mp = z[f"{sname}__opencv_calibration_multiplane__records"][i]
        if mp[14] == 1 and not feasible_record(mp):
            return "ill-posed"
    thrs = (ALGOS[algo],)
    if algo == "iterative_voter":
        # the thresholds the cascade visited: 0.5 (original_voter, voter), then 0.35, 0.2 until a camera came out
        branch = str(z[f"{sname}__{algo}__branch"][i])
        last = 0.5 if branch.startswith("ov_") else (float(branch.split("@")[1]) if "@" in branch else 0.2)
        thrs = tuple(t for t in (0.5, 0.35, 0.2) if t >= last)
    if any(degenerate_goal_view(z[f"{sname}__preds"][i], t) for t in thrs):
        return "ill-posed"
    if bool(z[f"{sname}__{algo}__minimal"][i]):
        return "minimal"
    if bool(z[f"{sname}__{algo}__ransac"][i]):
        return "ransac"
    return "exact"


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    scale = np.maximum(np.abs(ref[:14]), 1.0)
    return float(np.max(np.abs(got[:14] - ref[:14]) / scale))


def compare(solver: Callable[[np.ndarray, str, float], np.ndarray], algos=None, sets=SETS) -> Dict[str, dict]:
    """Runs ``solver(preds, algo, thr) -> (B,16)`` over the golden cases; returns per class
    {'n', 'decision_agree', 'param_agree', 'cameras', 'max_err', 'failures': [...]} and, under the key
    'per_algorithm', {algo: {'n', 'asserted', 'agree', 'cameras'}} over the asserted classes."""
    z = np.load(GOLDEN)
    stats: Dict[str, dict] = {}
    per_algo: Dict[str, dict] = {}
    for sname in sets:
        preds = np.ascontiguousarray(z[f"{sname}__preds"])
        for algo in (algos or ALGOS):
            got = solver(preds, algo, ALGOS[algo])
            ref = z[f"{sname}__{algo}__records"]
            for i in range(preds.shape[0]):
                cls = classify(z, sname, algo, i)
                s = stats.setdefault(cls, dict(n=0, decision_agree=0, param_agree=0, cameras=0, max_err=0.0, failures=[]))
                pa = per_algo.setdefault(algo, dict(n=0, asserted=0, agree=0, cameras=0))
                s["n"] += 1
                pa["n"] += 1
                pa["asserted"] += cls in ASSERTED
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
                    s["cameras"] += 1
                    pa["cameras"] += cls in ASSERTED
                s["param_agree"] += 1
                pa["agree"] += cls in ASSERTED
    stats["per_algorithm"] = per_algo
    return stats


def report(stats: Dict[str, dict]) -> str:
    """The per-class / per-algorithm table (printed by smoke() and the GPU test)."""
    lines = ["camera parity vs the reference's stored outputs (tests/golden/camera_cases.npz), tolerance 1e-4:"]
    for cls in ASSERTED + ("unpinned", "unconverged", "ill-posed"):
        if cls not in stats:
            continue
        v = stats[cls]
        lines.append(f"  {cls:12s} {'asserted' if cls in ASSERTED else 'reported'}  cases {v['n']:4d}  decisions agree {v['decision_agree']:4d}"
                     f"  parameters agree {v['param_agree']:4d}  (cameras {v['cameras']:3d}, max err {v['max_err']:.1e})")
    for algo, v in stats.get("per_algorithm", {}).items():
        lines.append(f"  {algo:30s} asserted {v['asserted']:3d}/{v['n']:3d}  agree {v['agree']:3d}  of which cameras {v['cameras']:3d}")
    return "\n".join(lines)


def assert_parity(stats: Dict[str, dict]) -> None:
    for cls in ASSERTED:
        v = stats[cls]
        assert not v["failures"], (cls, v["failures"][:5])
        assert v["max_err"] < TOL
    assert stats["exact"]["n"] > 700 and stats["minimal"]["n"] > 50 and stats["ransac"]["n"] > 15
    pa = stats["per_algorithm"]
    assert pa["iterative_voter"]["asserted"] >= 194 and pa["voter"]["cameras"] >= 30, pa


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
