"""GPU parity of the official metric (csrc/evaluate.cu through the C ABI) against the stored outputs of the
UNMODIFIED reference (tests/golden/evaluate_cases.npz) and against the oracle on a fresh end-to-end batch
(predictions -> CameraCreator -> metric), plus the API mirrors of baseline/evaluate_camera.py."""
import os

import numpy as np
import pytest
import torch

from oracle import camera_ref, evaluate_ref as O
from soccernet_calibration_sportlight_b200 import evaluate_camera as E, pitch, prediction
from soccernet_calibration_sportlight_b200.camera import Camera
from tests import camera_inputs as CI
from tests.test_oracle_evaluate import NC, case_inputs, load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_golden_metric_exact(golden_dir):
    """Integer outputs (confusion matrices, labelling, counts) bit-exact, accuracy exact, distance sums to 1e-9
    relative; the polylines equal the reference's to 1e-9 pixels."""
    z = load(golden_dir)
    rec = torch.from_numpy(z["records"]).to(DEV)
    res, (poly, pcnt, names) = E.evaluate_records(rec, torch.from_numpy(z["gt_pts"]).to(DEV), torch.from_numpy(z["gt_count"]).to(DEV), 5)
    ref = z["results"]
    for i, r in enumerate(res):
        assert r["valid"] == 1
        assert r["accuracy"] == ref[i, 0], i
        assert np.array_equal(r["confusion"], ref[i, 1:5]), i
        assert r["l2_count"] == ref[i, 6] and r["labelling"] == ref[i, 7], i
        assert abs(r["l2_sum"] - ref[i, 5]) <= 1e-9 * max(1.0, ref[i, 5]), i
        for c in range(NC):
            assert r["touched"][c] == ref[i, 8 + 5 * c + 4]
            assert np.array_equal(r["per_class"][c], ref[i, 8 + 5 * c: 8 + 5 * c + 4]), (i, c)
    poly, pcnt = poly.cpu().numpy(), pcnt.cpu().numpy()
    for i in range(6):
        for s, k in enumerate(names):
            key = f"poly__{i}__{pitch.LINES_CLASSES.index(k)}"
            assert (pcnt[i, s] > 0) == (key in z.files)
            if pcnt[i, s] > 0:
                assert pcnt[i, s] == len(z[key]) and np.abs(poly[i, s, :pcnt[i, s]] - z[key]).max() < 1e-9


def test_api_mirrors_vs_oracle(golden_dir):
    z = load(golden_dir)
    cam_o, annot = case_inputs(z, 2)
    cam = Camera(960, 540)
    cam.position, cam.rotation = cam_o.position, cam_o.rotation
    cam.xfocal_length, cam.yfocal_length = cam_o.xfocal_length, cam_o.yfocal_length
    pl = E.get_polylines(cam, 960, 540, sampling_factor=0.9)
    pl_o = O.get_polylines(cam_o, 960, 540, 0.9)
    assert pl.keys() == pl_o.keys()
    for k in pl:
        assert np.abs(np.array([[p["x"], p["y"]] for p in pl[k]]) - np.array([[p["x"], p["y"]] for p in pl_o[k]])).max() < 1e-9
    js = cam.to_json_parameters()                                   # the json route of get_polylines (camera.py:177-218)
    pl2 = E.get_polylines(js, 960, 540, sampling_factor=0.9)
    assert pl2.keys() == pl.keys()
    conf, per_class, errs = E.evaluate_camera_prediction(pl_o, annot, 5)
    conf_o, per_class_o, errs_o = O.evaluate_camera_prediction(pl_o, annot, 5)
    assert np.array_equal(conf, conf_o) and per_class.keys() == per_class_o.keys()
    assert all(np.array_equal(per_class[k], per_class_o[k]) for k in per_class)
    assert errs["count"] == sum(len(v) for v in errs_o.values())
    assert E.mirror_labels(annot).keys() == O.mirror_labels(annot).keys()
    with pytest.raises(Exception):
        E.evaluate_records(torch.zeros((1, 16), dtype=torch.float64), *E.pack_annotations([{}], "cpu"))   # host tensors: no CPU fallback


def test_evalai_metric_end_to_end_vs_oracle():
    """predictions -> CameraCreator (CUDA solve) -> metric kernel, aggregated like EvalAImetric, against the oracle
    path (cv2 solve + numpy metric) on frames whose reference outcome is reproducible."""
    kw = {k: v for k, v in camera_ref.MAKE_SUBMIT_KWARGS.items() if k not in ("algorithm", "conf_thresh")}
    preds = CI.clean_predictions(12, seed=41)
    rng = np.random.default_rng(5)
    creator_o = camera_ref.make_submit_creator()
    annots, keep = [], []
    for i in range(12):
        cam_o = creator_o(preds[i], None)
        keep.append(bool(creator_o.pinned))
        vis = O.get_polylines(cam_o, 960, 540, 0.9) if cam_o is not None else {}
        annots.append({k: [{"x": p["x"] + rng.normal(0, 2.0), "y": p["y"] + rng.normal(0, 2.0)} for p in v[::max(1, len(v) // 3)][:4]]
                       for k, v in list(vis.items())[::2]})
    mine = prediction.CameraCreator(pitch.PITCH_POINTS, conf_thresh=0.5, algorithm="iterative_voter", **kw)
    metric = E.EvalAImetric(mine, threshold=5)
    metric.update({"prediction": torch.from_numpy(preds), "raw_annots": annots, "img_name": [None] * 12})
    m = metric.epoch_metrics()
    # oracle aggregation
    tot = miss = n_acc = 0
    acc_sum = tp = n_prec = n_rec = l2 = 0.0
    n_l2 = 0
    for i in range(12):
        tot += 1
        res = O.evaluate_frame(creator_o(preds[i], None), annots[i], 5)
        if res is None:
            miss += 1
            continue
        a, conf, _, errs = res
        acc_sum += a; n_acc += 1
        tp += conf[0, 0]; n_prec += conf[0, :].sum(); n_rec += conf[0, 0] + conf[1, 0]
        n_l2 += sum(len(v) for v in errs.values()); l2 += sum(sum(v) for v in errs.values())
    if all(keep):
        assert m["completeness"] == (tot - miss) / tot
        assert abs(m["eval_accuracy"] - acc_sum / max(n_acc, 1)) < 1e-6
        assert abs(m["eval_precision"] - tp / max(n_prec, 1)) < 1e-6 and abs(m["eval_recall"] - tp / max(n_rec, 1)) < 1e-6
        assert abs(m["l2_reprojection"] - l2 / max(n_l2, 1)) < 1e-4
    assert 0.0 <= m["evalai"] <= 1.0 and metric.total_frames == 12
    print("\nEvalAI metric on 12 synthetic frames:", {k: round(float(v), 4) for k, v in m.items()})
