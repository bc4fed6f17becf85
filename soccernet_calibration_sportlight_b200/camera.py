"""``Camera`` with the reference's attributes and methods (baseline/camera.py:77-426) for
the calibration path.  The two optimisation entry points - ``solve_pnp`` and
``refine_camera`` - run on the GPU through the C ABI (``cal_pnp_solve`` / ``cal_pnp_refine``);
the remaining members are the small closed-form accessors of the reference (projection of a
point, pan/tilt/roll, JSON dictionary)."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def pan_tilt_roll_to_orientation(pan, tilt, roll):
    """camera.py:7-28."""
    rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    rx = np.array([[1, 0, 0], [0, np.cos(tilt), -np.sin(tilt)], [0, np.sin(tilt), np.cos(tilt)]])
    return np.dot(rz(pan), np.dot(rx, rz(roll)))


def rotation_matrix_to_pan_tilt_roll(rotation):
    """camera.py:31-58: ZXZ angles, of the two solutions the one with the smaller |roll|."""
    o = np.transpose(rotation)
    first_tilt = np.arccos(o[2, 2])
    sols = []
    for tilt in (first_tilt, -first_tilt):
        s = 1.0 if np.sin(tilt) > 0.0 else -1.0
        sols.append((np.arctan2(s * o[0, 2], s * -o[1, 2]), tilt, np.arctan2(s * o[2, 0], s * o[2, 1])))
    return sols[0] if np.fabs(sols[0][2]) < np.fabs(sols[1][2]) else sols[1]


def rodrigues(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> rotation vector (what cv.Rodrigues returns for a rotation matrix)."""
    R = np.asarray(R, dtype=np.float64)
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((r ** 2).sum() * 0.25)
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        v = np.sqrt(np.maximum((np.diag(R) + 1) * 0.5, 0))
        v[1] *= -1.0 if R[0, 1] < 0 else 1.0
        v[2] *= -1.0 if R[0, 2] < 0 else 1.0
        if abs(v[0]) < abs(v[1]) and abs(v[0]) < abs(v[2]) and (R[1, 2] > 0) != (v[1] * v[2] > 0):
            v[2] = -v[2]
        return v * (theta / np.linalg.norm(v))
    return r * (0.5 * theta / s)


def rotation_from_rodrigues(w: np.ndarray) -> np.ndarray:
    w = np.asarray(w, dtype=np.float64).reshape(3)
    th = np.linalg.norm(w)
    if th < 1e-14:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


class Camera:
    def __init__(self, iwidth=960, iheight=540):
        self.position = np.zeros(3)
        self.rotation = np.eye(3)
        self.calibration = np.eye(3)
        self.radial_distortion = np.zeros(6)
        self.thin_prism_disto = np.zeros(4)
        self.tangential_disto = np.zeros(2)
        self.image_width = iwidth
        self.image_height = iheight
        self.xfocal_length = 1
        self.yfocal_length = 1
        self.principal_point = (self.image_width / 2, self.image_height / 2)
        self.device = "cuda:0"        # where solve_pnp / refine_camera run

    # -- optimisation entry points: CUDA --------------------------------------------------
    def _pnp(self, point_matches: Sequence, refine: bool) -> bool:
        import torch
        from . import ops
        dev = torch.device(self.device)
        obj = torch.tensor(np.array([pt[0] for pt in point_matches], dtype=np.float64), device=dev)
        img = torch.tensor(np.array([pt[1] for pt in point_matches], dtype=np.float64), device=dev)
        K = torch.tensor(np.asarray(self.calibration, dtype=np.float64), device=dev)
        rvec = torch.tensor(rodrigues(self.rotation), device=dev)
        tvec = torch.tensor(-np.asarray(self.rotation) @ np.asarray(self.position), device=dev)
        ok = ops.pnp(obj.contiguous(), img.contiguous(), K.contiguous(), rvec, tvec, refine)
        if ok:
            self.rotation = rotation_from_rodrigues(rvec.cpu().numpy())
            self.position = -np.transpose(self.rotation) @ tvec.cpu().numpy()
        return ok

    def solve_pnp(self, point_matches):
        """camera.py:92-103: rotation and position for the current calibration matrix."""
        self._pnp(point_matches, refine=False)

    def refine_camera(self, pointMatches):
        """camera.py:105-119: Levenberg-Marquardt refinement of rotation and position."""
        self._pnp(pointMatches, refine=True)

    # -- closed-form members --------------------------------------------------------------
    def to_json_parameters(self):
        """camera.py:156-175."""
        pan, tilt, roll = rotation_matrix_to_pan_tilt_roll(self.rotation)
        return {
            "pan_degrees": pan * 180. / np.pi,
            "tilt_degrees": tilt * 180. / np.pi,
            "roll_degrees": roll * 180. / np.pi,
            "position_meters": np.asarray(self.position).tolist(),
            "x_focal_length": self.xfocal_length,
            "y_focal_length": self.yfocal_length,
            "principal_point": [self.principal_point[0], self.principal_point[1]],
            "radial_distortion": self.radial_distortion.tolist(),
            "tangential_distortion": self.tangential_disto.tolist(),
            "thin_prism_distortion": self.thin_prism_disto.tolist(),
        }

    def from_json_parameters(self, calib_json_object):
        """camera.py:177-218."""
        d = calib_json_object
        self.principal_point = d["principal_point"]
        self.image_width = 2 * self.principal_point[0]
        self.image_height = 2 * self.principal_point[1]
        self.xfocal_length = d["x_focal_length"]
        self.yfocal_length = d["y_focal_length"]
        self.calibration = np.array([[self.xfocal_length, 0, self.principal_point[0]],
                                     [0, self.yfocal_length, self.principal_point[1]],
                                     [0, 0, 1]], dtype="float")
        pan, tilt, roll = (d[k] * np.pi / 180. for k in ("pan_degrees", "tilt_degrees", "roll_degrees"))
        self.rotation = np.transpose(pan_tilt_roll_to_orientation(pan, tilt, roll))
        self.position = np.array(d["position_meters"], dtype="float")
        self.radial_distortion = np.array(d["radial_distortion"], dtype="float")
        self.tangential_disto = np.array(d["tangential_distortion"], dtype="float")
        self.thin_prism_disto = np.array(d["thin_prism_distortion"], dtype="float")

    def distort(self, point):
        """camera.py:220-247: rational radial (k1..k3 over k4..k6), tangential and thin-prism terms on a
        point of the normalised image plane; the result is float32 as there.  With the all-zero
        coefficients of this path it only rounds the point to float32."""
        x, y = point[0], point[1]
        r2 = x * x + y * y
        radius = np.sqrt(r2)
        num = den = 1
        for i in range(3):
            num += self.radial_distortion[i] * radius ** (2 * (i + 1))
            den += self.radial_distortion[i + 3] * radius ** (2 * (i + 1))
        f = num / den
        t, q = self.tangential_disto, self.thin_prism_disto
        xd = x * f + 2 * t[0] * x * y + t[1] * (radius ** 2 + 2 * x ** 2) + q[0] * radius ** 2 + q[1] * radius ** 4
        yd = y * f + 2 * t[1] * x * y + t[0] * (radius ** 2 + 2 * y ** 2) + q[2] * radius ** 2 + q[3] * radius ** 4
        return np.array([xd, yd], dtype=np.float32)

    def estimate_calibration_matrix_from_plane_homography(self, homography):
        """camera.py:366-426 (Hartley & Zisserman, algorithm 8.2): the image of the absolute conic w from
        the two circular-point constraints of a plane homography plus zero skew, square pixels and the
        principal-point ratio; K = (chol(w)^T)^-1 normalised.  Only the focal lengths are kept, the
        principal point stays at the image centre.  Returns (ok, K) like the reference; on a
        non-positive-definite w nothing is changed and (False, identity) comes back.
        (csrc/solve_cascade.cuh k_from_homography is the same computation inside the batched solve.)"""
        h = np.reshape(np.asarray(homography, dtype=np.float64), 9)
        c1, c2 = h[[0, 3, 6]], h[[1, 4, 7]]                   # first two columns of H

        def quad(a, b):                                       # a^T w b in the unknowns (w11 w12 w22 w13 w23 w33)
            return [a[0] * b[0], a[0] * b[1] + a[1] * b[0], a[1] * b[1], a[0] * b[2] + a[2] * b[0],
                    a[1] * b[2] + a[2] * b[1], a[2] * b[2]]
        A = np.zeros((5, 6))
        A[0, 1] = 1.0                                         # zero skew
        A[1, 0], A[1, 2] = 1.0, -1.0                          # square pixels
        A[2, 3], A[2, 4] = self.principal_point[1] / self.principal_point[0], -1.0
        A[3] = quad(c1, c2)                                   # h1^T w h2 = 0
        A[4] = np.subtract(quad(c1, c1), quad(c2, c2))        # h1^T w h1 = h2^T w h2
        w = np.linalg.svd(A)[2][-1]
        W = np.array([[w[0], w[1], w[3]], [w[1], w[2], w[4]], [w[3], w[4], w[5]]]) / w[5]
        try:
            L = np.linalg.cholesky(W)
        except np.linalg.LinAlgError:
            return False, np.eye(3)
        K = np.linalg.inv(L.T)
        K /= K[2, 2]
        self.xfocal_length, self.yfocal_length = K[0, 0], K[1, 1]
        self.principal_point = (self.image_width / 2, self.image_height / 2)
        self.calibration = np.array([[self.xfocal_length, 0, self.principal_point[0]],
                                     [0, self.yfocal_length, self.principal_point[1]],
                                     [0, 0, 1]], dtype="float")
        return True, K

    def project_point(self, point3D, distort=True):
        """camera.py:249-268."""
        p = self.rotation @ np.transpose(np.asarray(point3D) - self.position)
        if p[2] <= 1e-3:
            return np.zeros(3)
        p = p / p[2]
        if distort:
            p = self.distort(p)
        return np.array([p[0] * self.xfocal_length + self.principal_point[0],
                         p[1] * self.yfocal_length + self.principal_point[1], 1])

    def projection_rmse(self, matched_points):
        """camera.py:270-277: mean L2 distance (not an RMS)."""
        obj = np.array([pt[0] for pt in matched_points])
        img = np.array([pt[1] for pt in matched_points])
        proj = np.stack([self.project_point(p)[:2] for p in obj], axis=0)
        return np.mean(np.linalg.norm(img - proj, ord=2.0, axis=-1))

    def scale_resolution(self, factor):
        """camera.py:279-295."""
        self.xfocal_length = self.xfocal_length * factor
        self.yfocal_length = self.yfocal_length * factor
        self.image_width = self.image_width * factor
        self.image_height = self.image_height * factor
        self.principal_point = (self.image_width / 2, self.image_height / 2)
        self.calibration = np.array([[self.xfocal_length, 0, self.principal_point[0]],
                                     [0, self.yfocal_length, self.principal_point[1]],
                                     [0, 0, 1]], dtype="float")
