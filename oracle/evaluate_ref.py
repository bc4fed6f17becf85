"""ORACLE (test infrastructure, never shipped or timed as product): numpy restatement of the official
camera-calibration metric - ``baseline/evaluate_camera.py:14-229`` (get_polylines, distance_to_polyline,
evaluate_camera_prediction), ``baseline/evaluate_extremities.py:12-34`` (distance, mirror_labels) and the
per-frame evaluator of ``src/models/hrnet/metrics.py:107-137``.

Pinned: ``tests/golden/make_golden_evaluate.py`` runs the UNMODIFIED reference functions and this file on
the same synthetic cameras and annotations and stores the reference's outputs in
``tests/golden/evaluate_cases.npz``; ``tests/test_oracle_evaluate.py`` requires this file to reproduce them.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from soccernet_calibration_sportlight_b200 import pitch as _pitch


def get_polylines(cam, width: int, height: int, sampling_factor: float = 0.2) -> Dict[str, List[Dict[str, float]]]:
    """evaluate_camera.py:14-107; ``cam`` is any object with the reference Camera's project_point."""
    projections: Dict[str, List[Dict[str, float]]] = {}
    sides = [np.array([1, 0, 0]), np.array([1, 0, -width + 1]), np.array([0, 1, 0]), np.array([0, 1, -height + 1])]

    def crossing(ext, prev):
        line = np.cross(ext, prev)
        cands, dists = [], []
        for side in sides:
            with np.errstate(all="ignore"):
                it = np.cross(line, side)
                it = it / it[2]
            if 0 <= it[0] < width and 0 <= it[1] < height:
                cands.append(it)
                dists.append(np.sqrt(np.sum(np.square(it - ext))))
        return cands[int(np.argmin(dists))] if cands else None

    for key, points in _pitch.sample_field_points(sampling_factor).items():
        plist: List[Dict[str, float]] = []
        in_img = False
        prev = np.zeros(3)
        for i, point in enumerate(points):
            ext = cam.project_point(point)
            if ext[2] < 1e-5:
                continue
            if 0 <= ext[0] < width and 0 <= ext[1] < height:
                if not in_img and i > 0:
                    it = crossing(ext, prev)
                    if it is not None:
                        plist.append({"x": it[0], "y": it[1]})
                plist.append({"x": ext[0], "y": ext[1]})
                in_img = True
            elif in_img:
                it = crossing(ext, prev)
                if it is not None:
                    plist.append({"x": it[0], "y": it[1]})
                in_img = False
            prev = ext
        if plist:
            projections[key] = plist
    return projections


def distance(p1, p2) -> float:
    """evaluate_extremities.py:12-21."""
    return float(np.sqrt(np.square(np.array([p1["x"], p1["y"]]) - np.array([p2["x"], p2["y"]])).sum()))


def distance_to_polyline(point, polyline) -> float:
    """evaluate_camera.py:110-160."""
    if 0 < len(polyline) < 2:
        return distance(point, polyline[0])
    pt = np.array([point["x"], point["y"], 1])
    out = []
    for i in range(len(polyline) - 1):
        o = np.array([polyline[i]["x"], polyline[i]["y"], 1])
        e = np.array([polyline[i + 1]["x"], polyline[i + 1]["y"], 1])
        with np.errstate(all="ignore"):
            line = np.cross(o, e)
            line = line / np.sqrt(np.square(line[0]) + np.square(line[1]))
            pr = np.cross(np.cross(np.array([line[0], line[1], 0]), pt), line)
            pr = pr / pr[2]
            v1, v2 = pr - o, e - o
            k = np.dot(v1, v2) / np.dot(v2, v2)
        if 0 < k < 1:
            out.append(np.sqrt(np.sum(np.square(pr - pt))))
        else:
            out.append(np.min([distance(point, polyline[i]), distance(point, polyline[i + 1])]))
    return float(np.min(out))


def evaluate_camera_prediction(projected, groundtruth, threshold):
    """evaluate_camera.py:163-229 -> (global 2x2 float32 confusion, per-class confusions, per-class distances)."""
    conf = np.zeros((2, 2), dtype=np.float32)
    per_class, errors = {}, {}
    det, gt = set(projected), set(groundtruth)
    for c in det - gt:
        per_class[c] = np.array([[0.0, 2.0 if "Circle" not in c else 9.0], [0.0, 0.0]])
        conf[0, 1] += 1
    for c in gt - det:
        per_class[c] = np.array([[0.0, 0.0], [len(groundtruth[c]), 0.0]])
        conf[1, 0] += 1
    for c in det - (det - gt):
        per_class[c] = np.zeros((2, 2))
        ok = 1
        for p in groundtruth[c]:
            d = distance_to_polyline(p, projected[c])
            if d < threshold:
                per_class[c][0, 0] += 1
            else:
                per_class[c][0, 1] += 1
                ok = 0
            errors.setdefault(c, []).append(d)
        if ok:
            conf[0, 0] += 1
        else:
            conf[0, 1] += 1
    return conf, per_class, errors


def mirror_labels(lines):
    """evaluate_extremities.py:24-34."""
    return {_pitch.symmetric_class(k): v for k, v in lines.items()}


def evaluate_frame(cam, annot, threshold=5, img_size=(960, 540)) -> Optional[Tuple[float, np.ndarray, dict, dict]]:
    """Evaluator.__call__ after pred2cam (metrics.py:109-137): the better of the annotated and the mirrored labelling."""
    if cam is None:
        return None
    pred = get_polylines(cam, img_size[0], img_size[1], sampling_factor=0.9)
    r1 = evaluate_camera_prediction(pred, annot, threshold)
    r2 = evaluate_camera_prediction(pred, mirror_labels(annot), threshold)
    a1 = r1[0][0, 0] / r1[0].sum() if r1[0].sum() > 0 else 0.0
    a2 = r2[0][0, 0] / r2[0].sum() if r2[0].sum() > 0 else 0.0
    return (a1,) + r1 if a1 > a2 else (a2,) + r2
