"""ORACLE (test infrastructure, never shipped or timed as product): CPU restatement of the
reference's camera solve - ``src/models/hrnet/prediction.py`` (CameraCreator and its
candidate-camera helpers) on top of ``baseline/camera.py`` (Camera) - calling the same
third-party routines the reference calls (opencv-python: calibrateCamera, solvePnPRansac,
solvePnPRefineLM, findHomography, Rodrigues; the reference pins 4.7.0.72,
requirements.txt:5, this image ships 4.13.0) and numpy.linalg.

Pinned: ``tests/golden/make_golden_camera.py`` runs the UNMODIFIED reference classes and
this restatement on the same synthetic keypoint sets and stores the reference's outputs in
``tests/golden/camera_cases.npz``; ``tests/test_oracle_camera.py`` requires this file to
reproduce them exactly (same cv2 build => bit-identical).

Differences in form, not in arithmetic: functions return ``(camera, branch)`` where
``branch`` names the heuristic that produced the camera (the reference only prints it), and
nothing is printed.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import cv2
import numpy as np

from soccernet_calibration_sportlight_b200 import pitch as _pitch

TOP = (0, 1, 24, 25)                                     # prediction.py:15
PLANES = {                                               # prediction.py:18-23
    "groundplane": [i for i in range(58) if i not in TOP],
    "goal_left": [0, 1, 2, 3, 6, 7, 10, 11, 12, 13],
    "goal_right": [18, 19, 22, 23, 24, 25, 26, 27, 28, 29],
}
KEEP = list(range(29)) + [40, 41, 42, 44, 45, 48, 51, 52, 55]   # prediction.py:25-26
IMG_SIZE = (960, 540)                                    # prediction.py:24
WORLD = _pitch.keypoint_world_table()                    # (57,3) == PITCH_POINTS[INTERSECTON_TO_PITCH_POINTS[i]]

_FIX_ALL = (cv2.CALIB_FIX_PRINCIPAL_POINT | cv2.CALIB_FIX_ASPECT_RATIO | cv2.CALIB_FIX_TANGENT_DIST
            | cv2.CALIB_FIX_S1_S2_S3_S4 | cv2.CALIB_FIX_TAUX_TAUY | cv2.CALIB_FIX_K1 | cv2.CALIB_FIX_K2
            | cv2.CALIB_FIX_K3 | cv2.CALIB_FIX_K4 | cv2.CALIB_FIX_K5 | cv2.CALIB_FIX_K6)   # prediction.py:216-222
_FIX_SIMPLE = (cv2.CALIB_FIX_PRINCIPAL_POINT | cv2.CALIB_FIX_ASPECT_RATIO | cv2.CALIB_FIX_K1 | cv2.CALIB_FIX_K2
               | cv2.CALIB_FIX_K3 | cv2.CALIB_FIX_K4 | cv2.CALIB_FIX_TANGENT_DIST)         # prediction.py:152-155


def plane_coords(plane: str, p: np.ndarray) -> np.ndarray:
    """sets_transforms (prediction.py:29-41): goal planes are made z=0 by (x,y,z)->(y,z,0)."""
    if plane == "groundplane":
        return p
    return np.array([p[1], p[2], 0.0])


# ------------------------------------------------------------------------------- Camera
class CameraRef:
    """baseline/camera.py:77-426, the members the calibration path touches."""

    # bookkeeping of this restatement only (golden-vector generation): when set, every
    # refine_camera is repeated from its own result with a tight stopping rule; a result that still
    # moves means OpenCV's LM stopped (step < 1e-5) before reaching a stationary point, so the
    # reference's answer there is "where its solver happened to stop", not a minimiser
    probe_convergence = False

    def __init__(self, iwidth=960, iheight=540):
        self.position = np.zeros(3)
        self.rotation = np.eye(3)
        self.calibration = np.eye(3)
        self.radial_distortion = np.zeros(6)
        self.thin_prism_disto = np.zeros(4)
        self.tangential_disto = np.zeros(2)
        self.image_width, self.image_height = iwidth, iheight
        self.xfocal_length = 1
        self.yfocal_length = 1
        self.principal_point = (iwidth / 2, iheight / 2)
        # bookkeeping of this restatement only: set when solvePnPRansac reported failure, in which
        # case the reference goes on with the UNINITIALISED rvec/tvec OpenCV hands back (it ignores
        # the return flag, camera.py:100-103) and its result is not reproducible run to run
        self.tainted = False
        # ... and when solvePnPRansac was handed exactly its minimal sample size (5 points, or 4 where
        # it switches to P3P): OpenCV then returns the minimal solver's answer directly (no RANSAC,
        # no iterative refit); on the coplanar pitch points that answer is typically the flipped
        # planar pose.  Deterministic, but outside what the CUDA solver restates (DESIGN.md,
        # "parity classes").
        self.minimal = False
        # ... and whenever solve_pnp ran at all: even a successful solvePnPRansac seeds its answer from
        # EPnP on random 5-point samples (OpenCV's internal RNG), and on the near-coplanar pitch
        # points that is often the flipped planar pose, which refine_camera then keeps.
        self.ransac = False
        self.unconverged = False      # see probe_convergence

    # camera.py:92-103
    def solve_pnp(self, matches):
        obj = np.array([m[0] for m in matches])
        img = np.array([m[1] for m in matches])
        ok, rvec, t, _ = cv2.solvePnPRansac(obj, img, self.calibration, None)
        if not ok:
            self.tainted = True
        if len(matches) in (4, 5):
            self.minimal = True
        self.ransac = True
        self.rotation, _ = cv2.Rodrigues(rvec)
        self.position = -self.rotation.T @ t.flatten()

    # camera.py:105-119
    def refine_camera(self, matches):
        rvec, _ = cv2.Rodrigues(self.rotation)
        obj = np.array([m[0] for m in matches])
        img = np.array([m[1] for m in matches])
        crit = (cv2.TERM_CRITERIA_MAX_ITER + cv2.TERM_CRITERIA_EPS, 20000, 0.00001)
        rvec, t = cv2.solvePnPRefineLM(obj, img, self.calibration, None, rvec, -self.rotation @ self.position, crit)
        if CameraRef.probe_convergence:
            tight = (cv2.TERM_CRITERIA_MAX_ITER + cv2.TERM_CRITERIA_EPS, 200000, 1e-13)
            r2, t2 = cv2.solvePnPRefineLM(obj, img, self.calibration, None, rvec.copy(), t.copy(), tight)
            moved = max(float(np.abs(r2 - rvec).max()), float((np.abs(t2 - t) / np.maximum(np.abs(t), 1.0)).max()))
            if not moved < 1e-6:
                self.unconverged = True
        self.rotation, _ = cv2.Rodrigues(rvec)
        self.position = -self.rotation.T @ t

    # camera.py:366-426
    def estimate_calibration_matrix_from_plane_homography(self, homography):
        h = np.reshape(homography, (9,))
        A = np.zeros((5, 6))
        A[0, 1] = 1.0
        A[1, 0], A[1, 2] = 1.0, -1.0
        A[2, 3], A[2, 4] = self.principal_point[1] / self.principal_point[0], -1.0
        A[3] = [h[0] * h[1], h[0] * h[4] + h[1] * h[3], h[3] * h[4], h[0] * h[7] + h[1] * h[6],
                h[3] * h[7] + h[4] * h[6], h[6] * h[7]]
        A[4] = [h[0] * h[0] - h[1] * h[1], 2 * h[0] * h[3] - 2 * h[1] * h[4], h[3] * h[3] - h[4] * h[4],
                2 * h[0] * h[6] - 2 * h[1] * h[7], 2 * h[3] * h[6] - 2 * h[4] * h[7], h[6] * h[6] - h[7] * h[7]]
        w = np.linalg.svd(A)[2][-1]
        W = np.array([[w[0], w[1], w[3]], [w[1], w[2], w[4]], [w[3], w[4], w[5]]]) / w[5]
        try:
            kt_inv = np.linalg.cholesky(W)
        except np.linalg.LinAlgError:
            return False, np.eye(3)
        K = np.linalg.inv(kt_inv.T)
        K /= K[2, 2]
        self.xfocal_length, self.yfocal_length = K[0, 0], K[1, 1]
        self.principal_point = (self.image_width / 2, self.image_height / 2)
        self.calibration = np.array([[self.xfocal_length, 0, self.principal_point[0]],
                                     [0, self.yfocal_length, self.principal_point[1]], [0, 0, 1]], dtype="float")
        return True, K

    # camera.py:249-268 (all distortion coefficients are zero on this path; distort() then only
    # casts the normalised point to float32, camera.py:247)
    def project_point(self, p3, distort=True):
        r = self.rotation @ np.transpose(p3 - self.position)
        if r[2] <= 1e-3:
            return np.zeros(3)
        r = r / r[2]
        if distort:
            r = np.array([r[0], r[1]], dtype=np.float32)
        return np.array([r[0] * self.xfocal_length + self.principal_point[0],
                         r[1] * self.yfocal_length + self.principal_point[1], 1])

    # camera.py:270-277 (mean L2, not RMS)
    def projection_rmse(self, matches):
        obj = np.array([m[0] for m in matches])
        img = np.array([m[1] for m in matches])
        proj = np.stack([self.project_point(p)[:2] for p in obj], axis=0)
        return np.mean(np.linalg.norm(img - proj, ord=2.0, axis=-1))

    # camera.py:31-58, 156-175
    def to_json_parameters(self):
        o = np.transpose(self.rotation)
        t1 = np.arccos(o[2, 2])
        sols = []
        for tilt in (t1, -t1):
            s = 1.0 if np.sin(tilt) > 0.0 else -1.0
            sols.append((np.arctan2(s * o[0, 2], s * -o[1, 2]), tilt, np.arctan2(s * o[2, 0], s * o[2, 1])))
        pan, tilt, roll = sols[0] if np.fabs(sols[0][2]) < np.fabs(sols[1][2]) else sols[1]
        return {"pan_degrees": pan * 180.0 / np.pi, "tilt_degrees": tilt * 180.0 / np.pi,
                "roll_degrees": roll * 180.0 / np.pi, "position_meters": self.position.tolist(),
                "x_focal_length": self.xfocal_length, "y_focal_length": self.yfocal_length,
                "principal_point": [self.principal_point[0], self.principal_point[1]],
                "radial_distortion": self.radial_distortion.tolist(),
                "tangential_distortion": self.tangential_disto.tolist(),
                "thin_prism_distortion": self.thin_prism_disto.tolist()}

    # camera.py:177-218 (+ pan_tilt_roll_to_orientation :7-28)
    def from_json_parameters(self, d):
        self.principal_point = d["principal_point"]
        self.image_width, self.image_height = 2 * self.principal_point[0], 2 * self.principal_point[1]
        self.xfocal_length, self.yfocal_length = d["x_focal_length"], d["y_focal_length"]
        self.calibration = np.array([[self.xfocal_length, 0, self.principal_point[0]],
                                     [0, self.yfocal_length, self.principal_point[1]], [0, 0, 1]], dtype="float")
        pan, tilt, roll = (d[k] * np.pi / 180.0 for k in ("pan_degrees", "tilt_degrees", "roll_degrees"))
        rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        rx = np.array([[1, 0, 0], [0, np.cos(tilt), -np.sin(tilt)], [0, np.sin(tilt), np.cos(tilt)]])
        self.rotation = np.transpose(np.dot(rz(pan), np.dot(rx, rz(roll))))
        self.position = np.array(d["position_meters"], dtype="float")
        return self


def _camera_from_calibration(mtx, rvec0, tvec0, size) -> CameraRef:
    """The block repeated at prediction.py:160-168, 228-237, 411-420, 625-633."""
    cam = CameraRef(*size)
    cam.calibration = mtx
    cam.xfocal_length, cam.yfocal_length = mtx[0, 0], mtx[1, 1]
    cam.principal_point = (size[0] / 2.0, size[1] / 2.0)
    cam.rotation, _ = cv2.Rodrigues(rvec0)
    cam.position = (-np.transpose(cam.rotation) @ tvec0).T[0]
    return cam


def matched(points: Dict[int, Tuple[float, float]]):
    """get_matched_points (prediction.py:464-466)."""
    return [(WORLD[i], points[i]) for i in points]


def feasible(mtx, pos) -> bool:
    """good_camera / is_good_camera (prediction.py:469-484)."""
    return bool(10 <= mtx[0, 0] <= 20000 and -250 < pos[0] < 250 and -250 < pos[1] < 250 and -100 < pos[2] < 0)


def _plane_lists(points, plane):
    ids = [i for i in PLANES[plane] if i in points]
    return ([plane_coords(plane, WORLD[i]) for i in ids], [points[i] for i in ids], ids)


def homography_camera(points) -> Optional[Tuple[CameraRef, float]]:
    """get_camera_from_homography (prediction.py:487-520)."""
    wp, cp, _ = _plane_lists(points, "groundplane")
    wp, cp = np.array(wp, dtype=np.float32), np.array(cp, dtype=np.float32)
    if cp.shape[0] >= 4:
        hom, _ = cv2.findHomography(wp, cp, cv2.RANSAC, 10)            # ellipse.py:496-498
        if hom is not None:
            cam = CameraRef()
            cam.estimate_calibration_matrix_from_plane_homography(hom)
            m = matched(points)
            cam.solve_pnp(m)
            cam.refine_camera(m)
            return cam, cam.projection_rmse(m)
    return None


def general_camera(world_lists, cam_lists, m, n_groundplane) -> Optional[Tuple[CameraRef, float]]:
    """get_camera_gen (prediction.py:609-640)."""
    if len(world_lists) > 0 and sum(len(s) for s in cam_lists) > 6:
        _, mtx, _, rv, tv = cv2.calibrateCamera(world_lists, cam_lists, IMG_SIZE, None, None, flags=_FIX_ALL)
        cam = _camera_from_calibration(mtx, rv[0], tv[0], IMG_SIZE)
        if n_groundplane < 6:
            cam.solve_pnp(m)
        if len(m) > 6:
            cam.refine_camera(m)
        return cam, cam.projection_rmse(m)
    return None


def all_points_camera(points) -> Optional[Tuple[CameraRef, float]]:
    """get_camera_all_points (prediction.py:523-555), including the duplicated-view behaviour:
    the append sits inside the per-id loop, so a plane's (final) list is passed once per id
    visited from its first detected id on; and ``len(points_dict['groundplane'])`` is the
    length of a 2-key dict (always 2 < 6 => solve_pnp always runs)."""
    world_s, cam_s = [], []
    for plane, ids in PLANES.items():
        wl, cl = [], []
        for i in ids:
            if i in points:
                wl.append(plane_coords(plane, WORLD[i]))
                cl.append(points[i])
            if len(cl) > 0:
                world_s.append(wl)
                cam_s.append(cl)
    world_l = [np.array(s, dtype=np.float32) for s in world_s if len(s) >= 6]
    cam_l = [np.array(s, dtype=np.float32) for s in cam_s if len(s) >= 6]
    try:
        return general_camera(world_l, cam_l, matched(points), 2)
    except Exception:
        return None


def reliable_points_camera(points):
    """prediction.py:558-562."""
    return all_points_camera({k: v for k, v in points.items() if k in KEEP})


def groundplane_points_camera(points):
    """prediction.py:565-569."""
    return all_points_camera({k: v for k, v in points.items() if k in PLANES["groundplane"]})


def accurate_points_camera(points, threshold=10.0):
    """get_camera_accurate_points (prediction.py:572-606)."""
    wp, cp, ids = _plane_lists(points, "groundplane")
    wp, cp = np.array(wp, dtype=np.float32), np.array(cp, dtype=np.float32)
    if cp.shape[0] >= 4:
        hom, _ = cv2.findHomography(wp, cp, cv2.RANSAC, threshold)
        if hom is not None:
            wh = np.concatenate((wp[:, :2], np.ones((wp.shape[0], 1))), axis=-1)
            proj = []
            for d in wh:
                q = hom @ d
                proj.append(q[:2] / q[2])
            err = np.linalg.norm(np.array(proj) - cp, 2, axis=1)
            sel = {}
            for k in range(len(err)):
                if err[k] < threshold:
                    sel[ids[k]] = points[ids[k]]
            for i in TOP:
                if i in points:
                    sel[i] = points[i]
            return all_points_camera(sel)
    return None


# ------------------------------------------------------------------------ CameraCreator
class CameraCreatorRef:
    """prediction.py:44-437.  ``line_points`` replaces the pickle plumbing of __init__
    (:104-124): a dict name -> {keypoint id: (x, y)} as built there."""

    ALGORITHMS = ("opencv_calibration", "opencv_calibration_multiplane", "voter", "iterative_voter", "original_voter")

    def __init__(self, pitch=None, img_size=(960, 540), conf_thresh=0.2, algorithm="opencv_calibration",
                 line_points: Optional[Dict[str, Dict[int, Tuple[float, float]]]] = None, **kwargs):
        assert algorithm in self.ALGORITHMS
        self.algorithm = algorithm
        self.conf_thresh = conf_thresh
        self.img_size = img_size
        self.lines_data = dict(line_points or {})
        for k, v in kwargs.items():
            setattr(self, k, v)
        self.branch = None
        self.pinned = True      # False when the outcome depended on a tainted camera (see CameraRef)
        self.minimal = False    # True when the outcome depended on OpenCV's 5-point EPnP shortcut
        self.ransac = False     # True when the outcome depended on any solvePnPRansac result
        self.unconverged = False  # (CameraRef.probe_convergence) a refine_camera the outcome depended on stopped early

    def __call__(self, pred, name=None):
        self.branch = None
        self.pinned = True
        self.minimal = False
        self.ransac = False
        self.unconverged = False
        try:
            return getattr(self, self.algorithm)(pred, name)
        except Exception:
            return None

    def _line_points(self, name):
        return self.lines_data.get(name, {}) if name is not None else {}

    def _select(self, pred, reliable_gate: bool):
        pts: Dict[int, Tuple[float, float]] = {}
        n_det = np.count_nonzero(pred[:, 2] > self.conf_thresh)
        for i in range(pred.shape[0]):
            if pred[i, 2] > self.conf_thresh and (not reliable_gate or n_det < self.reliable_thresh or i in KEEP):
                pts[i] = (float(pred[i, 0]), float(pred[i, 1]))
        return pts

    # prediction.py:138-170
    def opencv_calibration(self, pred, name=None):
        ids = [i for i in range(pred.shape[0]) if i not in TOP and pred[i, 2] > self.conf_thresh]
        if len(ids) <= 5:
            return None
        cp = [(float(pred[i, 0]), float(pred[i, 1])) for i in ids]
        _, mtx, _, rv, tv = cv2.calibrateCamera(np.array([[WORLD[i] for i in ids]], dtype=np.float32),
                                                np.array([cp], dtype=np.float32), self.img_size, None, None,
                                                flags=_FIX_SIMPLE)
        self.branch = "calibration"
        return _camera_from_calibration(mtx, rv[0], tv[0], self.img_size)

    def _views(self, pts):
        world_s, cam_s, per_plane = [], [], {}
        for plane in PLANES:
            wl, cl, _ = _plane_lists(pts, plane)
            per_plane[plane] = len(cl)
            if len(cl) > 0:
                world_s.append(wl)
                cam_s.append(cl)
        mpp = self.min_points_per_plane
        return ([np.array(s, dtype=np.float32) for s in world_s if len(s) >= mpp],
                [np.array(s, dtype=np.float32) for s in cam_s if len(s) >= mpp], per_plane)

    # prediction.py:172-243
    def opencv_calibration_multiplane(self, pred, name=None):
        pts = self._select(pred, True)
        for i, p in self._line_points(name).items():
            if i not in pts and len(pts) <= self.min_points:
                pts[i] = p
        world_l, cam_l, _ = self._views(pts)
        if len(cam_l) > 0 and len(pts) > self.min_points:
            _, mtx, _, rv, tv = cv2.calibrateCamera(world_l, cam_l, self.img_size, None, None, flags=_FIX_ALL)
            if mtx[0, 0] > self.min_focal_length:
                cam = _camera_from_calibration(mtx, rv[0], tv[0], self.img_size)
                self.branch = "multiplane"
                if len(pts) > self.min_points_for_refinement:
                    cam.refine_camera(matched(pts))
                    self.unconverged |= cam.unconverged
                    self.branch = "multiplane+refine"
                return cam
        return None

    # prediction.py:339-437
    def original_voter(self, pred, name=None):
        pts = self._select(pred, True)
        n_ground = sum(1 for i in pts if i not in TOP)
        for i, p in self._line_points(name).items():
            if i not in pts and (n_ground < self.min_points_per_plane or
                                 (0 <= p[0] <= self.img_size[0] and 0 <= p[1] <= self.img_size[1])):
                pts[i] = p
        m = matched(pts)
        hom = homography_camera(pts)
        world_l, cam_l, per_plane = self._views(pts)
        cam = None
        if len(cam_l) > 0 and len(pts) > self.min_points:
            _, mtx, _, rv, tv = cv2.calibrateCamera(world_l, cam_l, self.img_size, None, None, flags=_FIX_ALL)
            cam = _camera_from_calibration(mtx, rv[0], tv[0], self.img_size)
            self.branch = "ov_calibration"
            if per_plane["groundplane"] < self.min_points_per_plane:
                cam.solve_pnp(m)
                self.branch = "ov_calibration+pnp"
                self.pinned &= not cam.tainted
                self.minimal |= cam.minimal
                self.ransac |= cam.ransac
                self.unconverged |= cam.unconverged
            if not feasible(cam.calibration, cam.position):
                cam = None
            elif len(pts) > self.min_points_for_refinement:
                cam.refine_camera(m)
                self.unconverged |= cam.unconverged
                self.branch += "+refine"
        if cam is None and hom is not None:
            self.pinned &= not hom[0].tainted
            self.minimal |= hom[0].minimal
            self.ransac |= hom[0].ransac
            self.unconverged |= hom[0].unconverged
            if hom[1] < 26:
                cam = hom[0]
                self.branch = "ov_homography"
        if cam is None:
            self.branch = None
        return cam

    # prediction.py:259-330
    def voter(self, pred, name=None):
        pts = self._select(pred, False)
        for i, p in self._line_points(name).items():
            if i not in pts and sum(k in PLANES["groundplane"] for k in pts) < self.min_points_per_plane:
                pts[i] = p
        hom = homography_camera(pts)
        cand_all = all_points_camera(pts)
        cand_rel = reliable_points_camera(pts)
        cand_acc = accurate_points_camera(pts, 5.0)
        cand_gnd = groundplane_points_camera(pts)
        cams = []
        for cand, tag in ((cand_rel, "camera_rel"), (cand_acc, "camera_acc"), (cand_all, "cam_all"),
                          (cand_gnd, "cam_ground")):
            if cand is not None:
                self.pinned &= not cand[0].tainted
                self.minimal |= cand[0].minimal
                self.ransac |= cand[0].ransac
                self.unconverged |= cand[0].unconverged
            if cand is not None and feasible(cand[0].calibration, cand[0].position):
                cams.append((cand[0], cand[1], tag))
        cam = None
        if cams:
            best = max(cams, key=lambda c: (c[2] == "camera_rel" and c[1] < self.max_rmse_rel, 1 / c[1]))
            if best[1] < self.max_rmse:
                cam = best[0]
                self.branch = "voter_" + best[2]
        if cam is None and hom is not None:
            self.pinned &= not hom[0].tainted
            self.minimal |= hom[0].minimal
            self.ransac |= hom[0].ransac
            self.unconverged |= hom[0].unconverged
            if hom[1] < self.max_rmse:
                cam = hom[0]
                self.branch = "voter_homography"
        return cam

    # prediction.py:245-257
    def iterative_voter(self, pred, name=None):
        self.conf_thresh = 0.5
        try:
            cam = self.original_voter(pred, name)
            if cam is not None:
                return cam
        except Exception:
            pass
        for p in self.conf_threshs:
            self.conf_thresh = p
            cam = self.voter(pred, name)
            if cam is not None:
                self.branch += f"@{p}"
                return cam
        self.branch = None
        return None


MAKE_SUBMIT_KWARGS = dict(conf_thresh=0.5, conf_threshs=[0.5, 0.35, 0.2], algorithm="iterative_voter",
                          max_rmse=55.0, max_rmse_rel=5.0, min_points=5, min_focal_length=10.0,
                          min_points_per_plane=6, min_points_for_refinement=6, reliable_thresh=57)  # make_submit.py:45-50


def make_submit_creator(**over) -> CameraCreatorRef:
    kw = dict(MAKE_SUBMIT_KWARGS)
    kw.update(over)
    return CameraCreatorRef(**kw)


def solve(creator: CameraCreatorRef, pred: np.ndarray, line_kp: Optional[Dict[int, Tuple[float, float]]] = None):
    """One frame through ``CameraCreator.__call__``; ``line_kp`` = keypoints from line intersections."""
    name = None
    if line_kp:
        name = "frame"
        creator.lines_data = {name: line_kp}
    return creator(pred, name)


def line_keypoints(peaks: np.ndarray, prob_thre: float = 0.0) -> List[Dict[int, Tuple[float, float]]]:
    """(B,23,2,3) decoded+scaled peaks -> per frame {keypoint id: (x, y)}: get_line_data
    (export_line_result.py:85-131, scale already applied by the transform) followed by the
    intersection step of CameraCreator.__init__ (prediction.py:110-124)."""
    from . import decode_ref
    out = []
    for b in range(peaks.shape[0]):
        lines, _ = decode_ref.get_line_data(peaks[b:b + 1], _pitch.LINE_CLS, scale=1, prob_thre=prob_thre)
        out.append(decode_ref.lines_to_keypoints(lines, _pitch.LINE_INTERSECTIONS))
    return out


def camera_record(cam: Optional[CameraRef]) -> np.ndarray:
    """(16,) float64: position(3), rotation(9 row-major), fx, fy, valid, 0 - the layout of
    CalCameraRecord's numeric part, for comparisons."""
    r = np.zeros(16)
    if cam is None:
        return r
    r[0:3] = np.asarray(cam.position, dtype=np.float64).reshape(3)
    r[3:12] = np.asarray(cam.rotation, dtype=np.float64).reshape(9)
    r[12], r[13], r[14] = cam.xfocal_length, cam.yfocal_length, 1.0
    return r
