"""Golden vectors for the camera solve, produced by the UNMODIFIED reference (build
container only):
    python tests/golden/make_golden_camera.py
Reference entry points run here:
  src/models/hrnet/prediction.py:44-437   CameraCreator (all five algorithms)
  baseline/camera.py:77-426               Camera
Writes tests/golden/camera_cases.npz: the synthetic (N,57,3) keypoint sets, and per
algorithm the reference's camera records (position, rotation, fx, fy, valid), the branch the
oracle restatement took and a `pinned` flag.  A frame is unpinned when its outcome went through
a cv2.solvePnPRansac call that reported failure: the reference ignores the flag and consumes
the uninitialised rvec/tvec OpenCV returns, so ITS OWN result changes from run to run there.
Asserts the restatement (oracle/camera_ref.py) reproduces the reference bit for bit on every
pinned frame.  Also written, for the unit tests of the restated OpenCV routines
(csrc/solve_pnp_cv.cuh): tests/golden/cv_calls.npz - every cv2.solvePnPRansac and
cv2.findHomography(RANSAC) call the reference made on these frames, inputs and outputs."""
import os
import sys

import cv2
import numpy as np

import refimport

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
refimport.setup()
from src.datatools.ellipse import PITCH_POINTS  # noqa: E402
from src.models.hrnet.prediction import CameraCreator  # noqa: E402

from oracle import camera_ref as O  # noqa: E402
from tests import camera_inputs as CI  # noqa: E402

O.CameraRef.probe_convergence = True
pnp_calls, hom_calls = {}, {}
_ransac, _findh = cv2.solvePnPRansac, cv2.findHomography


def _rec_ransac(obj, img, K, dist, *a, **k):
    r = _ransac(obj, img, K, dist, *a, **k)
    key = (obj.tobytes(), img.tobytes(), K.tobytes())
    if _recording[0] and key not in pnp_calls:
        pnp_calls[key] = (obj.copy(), img.copy(), K.copy(), bool(r[0]), np.array(r[1]).ravel().copy(), np.array(r[2]).ravel().copy(),
                          np.zeros(0, int) if r[3] is None else np.array(r[3]).ravel().copy())
    return r


def _rec_findh(src, dst, method=0, thr=3.0, *a, **k):
    r = _findh(src, dst, method, thr, *a, **k)
    key = (src.tobytes(), dst.tobytes(), float(thr))
    if _recording[0] and method == cv2.RANSAC and key not in hom_calls:
        hom_calls[key] = (src.copy(), dst.copy(), float(thr), None if r[0] is None else r[0].copy())
    return r


_recording = [False]
cv2.solvePnPRansac, cv2.findHomography = _rec_ransac, _rec_findh

KW = dict(O.MAKE_SUBMIT_KWARGS)
KW.pop("algorithm")
KW.pop("conf_thresh")

sets = {
    "noisy": CI.synthetic_predictions(96, seed=0),
    "clean": CI.clean_predictions(32, seed=1),
    "sparse": CI.synthetic_predictions(48, seed=2, drop=0.55, conf_lo=0.3),
    "wide": CI.synthetic_predictions(32, seed=3, wide=True, noise_px=0.7, outlier=0.0),
}
algos = {"iterative_voter": 0.5, "original_voter": 0.5, "opencv_calibration": 0.5,
         "opencv_calibration_multiplane": 0.5, "voter": 0.35}
out = {}
lines = []
for sname, preds in sets.items():
    out[f"{sname}__preds"] = preds
    for algo, thr in algos.items():
        ref = CameraCreator(PITCH_POINTS, conf_thresh=thr, algorithm=algo, **KW)
        mine = O.CameraCreatorRef(conf_thresh=thr, algorithm=algo, **KW)
        recs = np.zeros((preds.shape[0], 16))
        branches, pinned, minimal, ransac, unconv = [], [], [], [], []
        for i in range(preds.shape[0]):
            _recording[0] = False
            with refimport.quiet():
                c_ref = ref(preds[i], f"{sname}_{i}")
            _recording[0] = True
            c_mine = mine(preds[i], f"{sname}_{i}")
            unconv.append(bool(mine.unconverged))
            recs[i] = O.camera_record(c_ref)
            r2 = O.camera_record(c_mine)
            same = np.array_equal(recs[i], r2) or (np.isnan(recs[i]).any() and np.array_equal(
                np.isnan(recs[i]), np.isnan(r2)))
            pinned.append(bool(mine.pinned))
            minimal.append(bool(mine.minimal))
            ransac.append(bool(mine.ransac))
            if not mine.pinned:
                branches.append((mine.branch or "none") + "!")
                continue
            assert same, (sname, algo, i, recs[i], r2)
            if c_ref is not None:
                j1, j2 = c_ref.to_json_parameters(), c_mine.to_json_parameters()
                assert all(np.array_equal(np.asarray(j1[k], dtype=float), np.asarray(j2[k], dtype=float),
                                          equal_nan=True) for k in j1), (sname, algo, i)
            branches.append(mine.branch or "none")
        out[f"{sname}__{algo}__records"] = recs
        out[f"{sname}__{algo}__branch"] = np.array(branches)
        out[f"{sname}__{algo}__pinned"] = np.array(pinned)
        out[f"{sname}__{algo}__minimal"] = np.array(minimal)
        out[f"{sname}__{algo}__ransac"] = np.array(ransac)
        out[f"{sname}__{algo}__unconverged"] = np.array(unconv)
        uniq, cnt = np.unique(branches, return_counts=True)
        lines.append(f"{sname:7s} {algo:30s} n={preds.shape[0]} " + " ".join(f"{u}:{c}" for u, c in zip(uniq, cnt)))
np.savez_compressed(os.path.join(HERE, "camera_cases.npz"), **out)
# the recorded OpenCV calls, padded to 57 points
pc = list(pnp_calls.values())
cv_out = {"pnp_n": np.array([len(c[0]) for c in pc]), "pnp_K": np.stack([c[2] for c in pc]),
          "pnp_ok": np.array([c[3] for c in pc]), "pnp_rvec": np.stack([c[4] for c in pc]),
          "pnp_tvec": np.stack([c[5] for c in pc])}
obj = np.zeros((len(pc), 57, 3)); img = np.zeros((len(pc), 57, 2)); inl = np.zeros((len(pc), 57), bool)
for j, c in enumerate(pc):
    obj[j, :len(c[0])] = c[0]; img[j, :len(c[0])] = c[1]; inl[j, c[6]] = True
cv_out.update(pnp_obj=obj, pnp_img=img, pnp_inliers=inl)
hc = list(hom_calls.values())
src = np.zeros((len(hc), 57, 2), np.float32); dst = np.zeros((len(hc), 57, 2), np.float32)
for j, c in enumerate(hc):
    src[j, :len(c[0])] = c[0][:, :2]; dst[j, :len(c[0])] = c[1]
cv_out.update(hom_n=np.array([len(c[0]) for c in hc]), hom_thr=np.array([c[2] for c in hc]), hom_src=src, hom_dst=dst,
              hom_ok=np.array([c[3] is not None for c in hc]),
              hom_H=np.stack([np.zeros((3, 3)) if c[3] is None else c[3] for c in hc]))
np.savez_compressed(os.path.join(HERE, "cv_calls.npz"), **cv_out)
print(f"recorded {len(pc)} solvePnPRansac and {len(hc)} findHomography calls")
open(os.path.join(HERE, "camera_report.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
