"""The pitch / plane / line tables rebuilt in pitch.py against a dump of the reference's
own tables (tests/golden/tables.json, written by make_golden_tables.py from
src/datatools/ellipse.py:16-185, prediction.py:15-26, intersections.py:13-44, line.py:35-57)."""
import json
import os

import numpy as np

from soccernet_calibration_sportlight_b200 import pitch as P


def _tables(golden_dir):
    with open(os.path.join(golden_dir, "tables.json")) as f:
        return json.load(f)


def test_pitch_points_match_reference(golden_dir):
    t = _tables(golden_dir)
    assert set(t["pitch_points"]) == set(P.PITCH_POINTS)
    for k, v in t["pitch_points"].items():
        np.testing.assert_allclose(P.PITCH_POINTS[k], np.array(v), rtol=0, atol=1e-12, err_msg=k)


def test_keypoint_ids_and_sets(golden_dir):
    t = _tables(golden_dir)
    assert {int(k): v for k, v in t["id_to_name"].items()} == P.INTERSECTON_TO_PITCH_POINTS
    assert t["top_gates"] == P.TOP_GATES
    assert t["point_sets"] == P.POINT_SETS
    assert t["keep_points"] == P.KEEP_POINTS
    assert t["points_left"] == P.POINTS_LEFT and t["points_right"] == P.POINTS_RIGHT
    assert tuple(t["img_size"]) == P.IMG_SIZE


def test_line_tables(golden_dir):
    t = _tables(golden_dir)
    assert {int(k): tuple(v) for k, v in t["line_intersections"].items()} == P.LINE_INTERSECTIONS
    assert {int(k): v for k, v in t["line_cls"].items()} == P.LINE_CLS


def test_world_table_shape():
    w = P.keypoint_world_table()
    assert w.shape == (57, 3) and w.dtype == np.float64
    assert (w[P.TOP_GATES, 2] == -2.44).all()
    ground = [i for i in range(57) if i not in P.TOP_GATES]
    assert (w[ground, 2] == 0).all()
