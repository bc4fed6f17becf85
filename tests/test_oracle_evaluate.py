"""CPU suite for the official metric: oracle/evaluate_ref.py against the stored outputs of the UNMODIFIED
reference (tests/golden/evaluate_cases.npz, made by tests/golden/make_golden_evaluate.py), and the
evaluation-side pitch tables of the product (class list, mirrored classes, sampled pitch model)."""
import os

import numpy as np

from oracle import camera_ref, evaluate_ref as O
from soccernet_calibration_sportlight_b200 import pitch

NC = 28


def load(golden_dir):
    return np.load(os.path.join(golden_dir, "evaluate_cases.npz"))


def case_inputs(z, i):
    rec = z["records"][i]
    cam = camera_ref.CameraRef(960, 540)
    cam.position, cam.rotation = rec[0:3].copy(), rec[3:12].reshape(3, 3).copy()
    cam.xfocal_length, cam.yfocal_length = rec[12], rec[13]
    annot = {}
    for c in range(NC):
        n = int(z["gt_count"][i, c])
        if n >= 0:
            annot[pitch.LINES_CLASSES[c]] = [{"x": float(x), "y": float(y)} for x, y in z["gt_pts"][i, c, :n]]
    return cam, annot


def test_oracle_reproduces_reference_metric(golden_dir):
    z = load(golden_dir)
    n_hit = 0
    for i in range(len(z["records"])):
        cam, annot = case_inputs(z, i)
        acc, conf, per_class, errs = O.evaluate_frame(cam, annot, 5)
        ref = z["results"][i]
        assert acc == ref[0] and np.array_equal(conf.reshape(4), ref[1:5])
        assert sum(len(v) for v in errs.values()) == ref[6]
        assert abs(sum(sum(v) for v in errs.values()) - ref[5]) <= 1e-9 * max(1.0, ref[5])
        for c in range(NC):
            touched = ref[8 + 5 * c + 4] == 1
            assert touched == (pitch.LINES_CLASSES[c] in per_class)
            if touched:
                assert np.array_equal(per_class[pitch.LINES_CLASSES[c]].reshape(4), ref[8 + 5 * c: 8 + 5 * c + 4])
        n_hit += acc > 0.5
    assert n_hit >= 10


def test_oracle_polylines_match_reference(golden_dir):
    z = load(golden_dir)
    for i in range(6):
        cam, _ = case_inputs(z, i)
        pl = O.get_polylines(cam, 960, 540, 0.9)
        keys = sorted(int(k.split("__")[2]) for k in z.files if k.startswith(f"poly__{i}__"))
        assert sorted(pitch.LINES_CLASSES.index(k) for k in pl) == keys
        for c in keys:
            got = np.array([[p["x"], p["y"]] for p in pl[pitch.LINES_CLASSES[c]]])
            assert np.array_equal(got, z[f"poly__{i}__{c}"])


def test_evaluation_tables():
    assert len(pitch.LINES_CLASSES) == NC and pitch.LINES_CLASSES == sorted(pitch.LINES_CLASSES)
    for k in pitch.LINES_CLASSES:
        assert pitch.symmetric_class(pitch.symmetric_class(k)) == k
    assert pitch.symmetric_class("Goal left post left ") == "Goal right post left"
    assert pitch.symmetric_class("Big rect. left top") == "Big rect. right bottom"
    s = pitch.sample_field_points(0.9)
    assert len(s) == 26 and sum(len(v) for v in s.values()) == 1187 and len(s["Circle central"]) == 287
    assert np.allclose(np.linalg.norm(s["Circle central"][:, :2], axis=1), 9.15)
