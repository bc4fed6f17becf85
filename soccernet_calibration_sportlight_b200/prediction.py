"""``CameraCreator`` with the reference's constructor and call signature
(src/models/hrnet/prediction.py:44-136): keypoint predictions -> ``Camera`` or ``None``.
The whole heuristic cascade (all five algorithms, :138-437, and the candidate-camera
helpers :464-640) runs in the batched CUDA kernel ``cal_camera_solve`` - one thread block per
frame - so ``__call__`` is a batch of one and ``batch`` / ``batch_records`` are the native
entry points (new API layered on top of the reference's, not replacing it).

    creator = CameraCreator(PITCH_POINTS, conf_thresh=0.5, algorithm='iterative_voter', ...)
    cam = creator(pred_57x3, name)                  # reference call
    cams = creator.batch(preds_Bx57x3)              # list[Optional[Camera]]
    recs = creator.batch_records(preds_on_device)   # (B,16) fp64 device tensor, no host sync
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .camera import Camera
from .pitch import INTERSECTON_TO_PITCH_POINTS, LINE_CLS, LINE_INTERSECTIONS, PITCH_POINTS  # noqa: F401

ALGORITHMS = {"opencv_calibration": 0, "opencv_calibration_multiplane": 1, "original_voter": 2, "voter": 3,
              "iterative_voter": 4}                                     # prediction.py:90-96
BRANCH_NAMES = {0: "none", 1: "calibration", 2: "multiplane", 3: "ov_calibration", 4: "ov_calibration+pnp",
                5: "ov_homography", 6: "voter_camera_rel", 7: "voter_camera_acc", 8: "voter_cam_all",
                9: "voter_cam_ground", 10: "voter_homography"}
MAKE_SUBMIT_KWARGS = dict(conf_thresh=0.5, conf_threshs=[0.5, 0.35, 0.2], algorithm="iterative_voter",
                          max_rmse=55.0, max_rmse_rel=5.0, min_points=5, min_focal_length=10.0,
                          min_points_per_plane=6, min_points_for_refinement=6, reliable_thresh=57)  # make_submit.py:45-50
_DEFAULTS = dict(min_points=5, min_focal_length=10.0, min_points_per_plane=6, min_points_for_refinement=6,
                 reliable_thresh=57, conf_threshs=[0.5, 0.35, 0.2], max_rmse=55.0, max_rmse_rel=5.0)


def line_eq_intersection(line1: Tuple[float, float], line2: Tuple[float, float]) -> Optional[Tuple[float, float]]:
    """prediction.py:643-653."""
    k1, b1 = line1
    k2, b2 = line2
    if abs(k1 - k2) > 1e-4:
        x = (b2 - b1) / (k1 - k2)
        return (x, k1 * x + b1)
    return None


def branch_name(code: int) -> str:
    return BRANCH_NAMES.get(code & 15, "?") + ("+refine" if code & 16 else "")


class CameraCreator:
    def __init__(self, pitch: Dict[str, np.ndarray], img_size: Tuple[int, int] = (960, 540), conf_thresh: float = 0.2,
                 algorithm: str = "opencv_calibration", lines_file: Optional[str] = None, **kwargs):
        assert algorithm in ALGORITHMS, f"Should be one of: {list(ALGORITHMS.keys())}"
        self.algorithm = algorithm
        self.conf_thresh = conf_thresh
        self.pitch = pitch
        self.img_size = img_size
        self.lines_data: Dict[str, Dict[int, Tuple[float, float]]] = {}
        if lines_file is not None:
            assert os.path.exists(lines_file), f"{lines_file} does not exist"
            with open(lines_file, "rb") as f:
                lines_data = pickle.load(f)
            for img_name, entry in lines_data.items():
                # prediction.py:112 reads entry['lines'][0]; export_line_result.py:188 stores the dict
                # itself - accept both layouts
                pred = entry["lines"]
                if not isinstance(pred, dict):
                    pred = pred[0]
                pred = dict(pred)
                if "Goal left post left" in pred:
                    pred["Goal left post left "] = pred.pop("Goal left post left")
                points = {}
                for idx, pair in LINE_INTERSECTIONS.items():
                    if pair[0] in pred and pair[1] in pred:
                        p = line_eq_intersection(pred[pair[0]], pred[pair[1]])
                        if p is not None:
                            points[idx] = p
                if points:
                    self.lines_data[img_name] = points
        for key, value in _DEFAULTS.items():
            setattr(self, key, value)
        for key, value in kwargs.items():
            setattr(self, key, value)
        self.stat = {"n": 0, "frames_4": 0, "frames_4_6": 0, "frames_bad_cam": 0}
        self.device = "cuda:0"
        self.last_branch: Optional[str] = None

    # -- parameters for the kernel ----------------------------------------------------------
    def _params(self) -> "_lib.SolveParams":
        P = _lib.SolveParams()
        for i, name in INTERSECTON_TO_PITCH_POINTS.items():
            xyz = np.asarray(self.pitch[name], dtype=np.float64)
            for k in range(3):
                P.pitch_xyz[3 * i + k] = float(xyz[k])
        P.algorithm = ALGORITHMS[self.algorithm]
        P.img_w, P.img_h = int(self.img_size[0]), int(self.img_size[1])
        P.conf_thresh = float(self.conf_thresh)
        ths = list(self.conf_threshs)[:8]
        for i, t in enumerate(ths):
            P.conf_threshs[i] = float(t)
        P.n_conf_threshs = len(ths)
        P.min_points, P.min_points_per_plane = int(self.min_points), int(self.min_points_per_plane)
        P.min_points_for_refinement, P.reliable_thresh = int(self.min_points_for_refinement), int(self.reliable_thresh)
        P.min_focal_length, P.max_rmse, P.max_rmse_rel = float(self.min_focal_length), float(self.max_rmse), float(self.max_rmse_rel)
        return P

    def _get_points_from_lines(self, name: Optional[str] = None):
        """prediction.py:332-337."""
        if self.lines_data is not None and name is not None and name in self.lines_data:
            return self.lines_data[name]
        return {}

    # -- native batched entry points --------------------------------------------------------
    def batch_records(self, preds, line_pts=None):
        """preds (B,57,3) fp32 on the GPU (+ optional (B,57,2) fp64 line keypoints, NaN = absent)
        -> (B,16) fp64 records on the GPU; nothing is synchronised."""
        from . import ops
        return ops.camera_solve(preds.contiguous(), self._params(), line_pts)

    def line_points_device(self, peaks, prob_thre: float = 0.0):
        """(B,23,2,3) decoded line peaks in image pixels -> (B,57,2) fp64 line-intersection keypoints."""
        import torch
        from . import ops
        if not hasattr(self, "_pairs") or self._pairs[0].device != peaks.device:
            name_to_ch = {v: k for k, v in LINE_CLS.items()}
            a = [-1] * 57
            b = [-1] * 57
            for idx, (la, lb) in LINE_INTERSECTIONS.items():
                a[idx], b[idx] = name_to_ch[la], name_to_ch[lb]
            self._pairs = (torch.tensor(a, dtype=torch.int32, device=peaks.device),
                           torch.tensor(b, dtype=torch.int32, device=peaks.device))
        return ops.line_points(peaks.contiguous(), self._pairs[0], self._pairs[1], prob_thre)

    def cameras_from_records(self, records: np.ndarray) -> List[Optional[Camera]]:
        records = np.ascontiguousarray(records, dtype=np.float64)
        flags = records.view(np.int32).reshape(records.shape[0], 32)[:, 30:32]
        cams: List[Optional[Camera]] = []
        self.last_branches = [branch_name(int(f[1])) if f[0] else None for f in flags]
        for r, f in zip(records, flags):
            if not f[0]:
                cams.append(None)
                continue
            cam = Camera(*self.img_size) if (int(f[1]) & 15) not in (5, 10) else Camera()
            cam.position = r[0:3].copy()
            cam.rotation = r[3:12].reshape(3, 3).copy()
            cam.xfocal_length, cam.yfocal_length = float(r[12]), float(r[13])
            cam.principal_point = (cam.image_width / 2.0, cam.image_height / 2.0)
            if (int(f[1]) & 15) in (5, 10):      # homography camera: K carries principal_point (camera.py:419-425)
                c = cam.principal_point
            else:                                # calibrateCamera's mtx: OpenCV's fixed centre ((W-1)/2, (H-1)/2)
                c = ((self.img_size[0] - 1) * 0.5, (self.img_size[1] - 1) * 0.5)
            cam.calibration = np.array([[r[12], 0, c[0]], [0, r[13], c[1]], [0, 0, 1]], dtype=np.float64)
            cam.device = self.device
            cams.append(cam)
        return cams

    def batch(self, preds: np.ndarray, names: Optional[List[Optional[str]]] = None) -> List[Optional[Camera]]:
        """(B,57,3) predictions -> list of ``Camera`` / ``None`` (one kernel launch)."""
        import torch
        preds = np.ascontiguousarray(preds, dtype=np.float32)
        dev = torch.device(self.device)
        lp = None
        if names is not None and self.lines_data:
            arr = np.full((preds.shape[0], 57, 2), np.nan)
            for b, nm in enumerate(names):
                for idx, p in self._get_points_from_lines(nm).items():
                    arr[b, idx] = p
            lp = torch.from_numpy(arr).to(dev)
        rec = self.batch_records(torch.from_numpy(preds).to(dev), lp)
        return self.cameras_from_records(rec.cpu().numpy())

    # -- the reference's call ----------------------------------------------------------------
    def __call__(self, pred, name: Optional[str] = None) -> Optional[Camera]:
        """prediction.py:130-136: any failure inside the solve gives ``None``."""
        cam = None
        try:
            cam = self.batch(np.asarray(pred)[None], [name])[0]
            self.last_branch = self.last_branches[0]
        except _lib.CalError:
            raise                                  # a missing library / device is not a 'bad frame'
        except Exception as e:                     # noqa: BLE001 - reference semantics
            print(f"Camera initialization exc: {e}")
        return cam
