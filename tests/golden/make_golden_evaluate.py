"""Golden vectors for the official metric, produced by the UNMODIFIED reference (build container only):
    python tests/golden/make_golden_evaluate.py
Reference functions run here: baseline/evaluate_camera.py:14-229 (get_polylines, evaluate_camera_prediction),
baseline/evaluate_extremities.py:24-34 (mirror_labels), baseline/camera.py (Camera), and the per-frame logic
of src/models/hrnet/metrics.py:109-137 (restated inline: metrics.py imports argus).  Writes
tests/golden/evaluate_cases.npz: predicted cameras (records), packed annotations, the reference's per-frame
results and a few of its polylines.  Asserts oracle/evaluate_ref.py reproduces every number."""
import os
import sys

import numpy as np

import refimport

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
refimport.setup()
from baseline.camera import Camera  # noqa: E402
from baseline.evaluate_camera import evaluate_camera_prediction, get_polylines  # noqa: E402
from baseline.evaluate_extremities import mirror_labels  # noqa: E402
from baseline.soccerpitch import SoccerPitch  # noqa: E402

from oracle import evaluate_ref as O  # noqa: E402
from soccernet_calibration_sportlight_b200 import pitch  # noqa: E402
from tests import camera_inputs as CI  # noqa: E402

NC, MAXGT = 28, 12
rng = np.random.default_rng(17)
CLS = SoccerPitch.lines_classes


def make_cam(R, pos, f):
    c = Camera(960, 540)
    c.rotation, c.position = R.copy(), pos.copy()
    c.xfocal_length = c.yfocal_length = np.float64(f)      # as CameraCreator sets it (mtx[0, 0]); a Python float would make numpy 2 multiply in float32
    c.calibration = np.array([[f, 0, 480.0], [0, f, 270.0], [0, 0, 1.0]])
    return c


recs, gts, cnts, results, polys = [], [], [], [], {}
for case in range(40):
    R, pos, f = CI.random_camera(rng)
    true_cam = make_cam(R, pos, f)
    # annotation: a few points of each visible class of the TRUE camera (+ pixel noise), some classes dropped,
    # sometimes an 'unknown' class, sometimes left/right swapped (what the mirrored labelling is for)
    vis = get_polylines(true_cam, 960, 540, sampling_factor=0.9)
    annot = {}
    for k, pl in vis.items():
        if rng.random() < 0.15:
            continue
        n = int(rng.integers(2, 9)) if "Circle" in k else int(rng.integers(2, 4))
        idx = np.sort(rng.choice(len(pl), size=min(n, len(pl)), replace=False))
        annot[k] = [{"x": pl[i]["x"] + rng.normal(0, 1.5), "y": pl[i]["y"] + rng.normal(0, 1.5)} for i in idx]
    if rng.random() < 0.3:
        annot["Line unknown"] = [{"x": float(rng.uniform(0, 960)), "y": float(rng.uniform(0, 540))} for _ in range(2)]
    if rng.random() < 0.25:
        annot = mirror_labels(annot)
    if case % 9 == 0:
        annot[list(annot)[0]] = annot[list(annot)[0]][:1]                    # a one-point class
    # prediction: the true camera perturbed (small: mostly hits; large: misses)
    s = 0.0004 if case % 3 else 0.01
    import cv2
    dR = cv2.Rodrigues(rng.normal(0, s, 3))[0]
    pred_cam = make_cam(dR @ R, pos + rng.normal(0, 30 * s, 3), f * (1 + rng.normal(0, s)))
    pred = get_polylines(pred_cam, 960, 540, sampling_factor=0.9)
    c1 = evaluate_camera_prediction(pred, annot, 5)
    c2 = evaluate_camera_prediction(pred, mirror_labels(annot), 5)
    a1 = c1[0][0, 0] / c1[0].sum() if c1[0].sum() > 0 else 0.0
    a2 = c2[0][0, 0] / c2[0].sum() if c2[0].sum() > 0 else 0.0
    acc, (conf, per_class, errs), lab = (a1, c1, 0) if a1 > a2 else (a2, c2, 1)
    # the oracle restatement must agree exactly
    o = O.evaluate_frame(pred_cam, annot, 5)
    assert o[0] == acc and np.array_equal(o[1], conf) and o[2].keys() == per_class.keys()
    assert all(np.array_equal(o[2][k], per_class[k]) for k in per_class)
    assert o[3].keys() == errs.keys() and all(np.array_equal(o[3][k], errs[k]) for k in errs)
    op = O.get_polylines(pred_cam, 960, 540, 0.9)
    assert op.keys() == pred.keys() and all(op[k] == pred[k] for k in pred)
    rec = np.zeros(16)
    rec[0:3], rec[3:12], rec[12], rec[13] = pred_cam.position, pred_cam.rotation.reshape(9), pred_cam.xfocal_length, pred_cam.yfocal_length
    rec.view(np.int32)[30] = 1
    g, n = np.zeros((NC, MAXGT, 2)), np.full(NC, -1, np.int32)
    for k, v in annot.items():
        n[CLS.index(k)] = len(v)
        for j, p in enumerate(v):
            g[CLS.index(k), j] = (p["x"], p["y"])
    res = np.zeros(8 + NC * 5)
    res[0], res[1:5], res[5], res[6], res[7] = acc, conf.reshape(4), sum(sum(v) for v in errs.values()), sum(len(v) for v in errs.values()), lab
    for k, m in per_class.items():
        res[8 + CLS.index(k) * 5: 8 + CLS.index(k) * 5 + 4] = m.reshape(4)
        res[8 + CLS.index(k) * 5 + 4] = 1
    recs.append(rec); gts.append(g); cnts.append(n); results.append(res)
    if case < 6:
        for k, v in pred.items():
            polys[f"poly__{case}__{CLS.index(k)}"] = np.array([[p["x"], p["y"]] for p in v])
assert pitch.LINES_CLASSES == CLS
np.savez_compressed(os.path.join(HERE, "evaluate_cases.npz"), records=np.array(recs), gt_pts=np.array(gts), gt_count=np.array(cnts),
                    results=np.array(results), **polys)
r = np.array(results)
print("cases", len(recs), "mean accuracy", r[:, 0].mean(), "mirrored chosen", int(r[:, 7].sum()), "mean l2", (r[:, 5].sum() / r[:, 6].sum()))
