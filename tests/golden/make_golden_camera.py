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
pinned frame."""
import os
import sys

import numpy as np

import refimport

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
refimport.setup()
from src.datatools.ellipse import PITCH_POINTS  # noqa: E402
from src.models.hrnet.prediction import CameraCreator  # noqa: E402

from oracle import camera_ref as O  # noqa: E402
from tests import camera_inputs as CI  # noqa: E402

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
        branches, pinned, minimal, ransac = [], [], [], []
        for i in range(preds.shape[0]):
            with refimport.quiet():
                c_ref = ref(preds[i], f"{sname}_{i}")
            c_mine = mine(preds[i], f"{sname}_{i}")
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
        uniq, cnt = np.unique(branches, return_counts=True)
        lines.append(f"{sname:7s} {algo:30s} n={preds.shape[0]} " + " ".join(f"{u}:{c}" for u, c in zip(uniq, cnt)))
np.savez_compressed(os.path.join(HERE, "camera_cases.npz"), **out)
open(os.path.join(HERE, "camera_report.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
