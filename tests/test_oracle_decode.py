"""oracle/decode_ref.py against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden_decode.py; src/models/hrnet/transforms.py:224-239,
src/models/line/transforms.py:216-280).  Bar: indices bit-exact; keypoint confidence within
1 ulp of the reference's torch-CPU exp (the oracle defines exp as correctly rounded, see
decode_report.txt); line values bit-exact."""
import os

import numpy as np
import pytest

from oracle import decode_ref as O
from tests import inputs as I

KP_CASES = {
    "hashed_small": lambda: I.hashed_logp(1, (2, 58, 20, 24)),
    "hashed_ragged": lambda: I.hashed_logp(2, (1, 5, 7, 13)),
    "hashed_full": lambda: I.hashed_logp(3, (1, 58, 270, 480)),
    "gauss_small": lambda: I.gaussian_logp(4, 2, 58, 20, 24),
    "gauss_full": lambda: I.gaussian_logp(5, 1, 58, 270, 480),
    "gauss_720p": lambda: I.gaussian_logp(6, 1, 58, 360, 640),
    "target_log": None, "logsoftmax_rand": None,
}
LINE_CASES = {
    "tent_small_s3": lambda: I.two_peak_heat(7, 2, 23, 17, 30),
    "tent_small_s6": lambda: I.two_peak_heat(8, 1, 23, 17, 30),
    "tent_full_s3": lambda: I.two_peak_heat(9, 1, 23, 135, 240),
    "tent_ragged": lambda: I.two_peak_heat(10, 1, 3, 5, 11),
    "softmax_rand": None,
}


def kp_case(z, name):
    x = z[f"{name}__in"] if KP_CASES[name] is None else KP_CASES[name]()
    return x, tuple(int(v) for v in z[f"{name}__size"]), z[f"{name}__out"]


def line_case(z, name):
    x = z[f"{name}__in"] if LINE_CASES[name] is None else LINE_CASES[name]()
    return x, float(z[f"{name}__sigma"]), z[f"{name}__out"], z[f"{name}__scaled4"]


@pytest.fixture(scope="module")
def kp_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "decode_keypoints.npz"))


@pytest.fixture(scope="module")
def line_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "decode_lines.npz"))


@pytest.mark.parametrize("name", list(KP_CASES))
def test_keypoint_oracle_matches_reference(kp_golden, name):
    x, size, ref = kp_case(kp_golden, name)
    got = O.keypoint_decode_np(x, size)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.array_equal(got[..., :2], ref[..., :2])                       # indices: bit-exact
    ulp = np.abs(got[..., 2].view(np.int32) - ref[..., 2].view(np.int32)).max()
    assert ulp <= 1


@pytest.mark.parametrize("name", list(KP_CASES))
def test_keypoint_torch_flavour_is_literal(kp_golden, name):
    import torch
    x, size, ref = kp_case(kp_golden, name)
    got = O.keypoint_decode_torch(torch.from_numpy(x), size).numpy()
    assert np.array_equal(got[..., :2], ref[..., :2])
    assert np.abs(got[..., 2].view(np.int32) - ref[..., 2].view(np.int32)).max() <= 1


@pytest.mark.parametrize("name", list(LINE_CASES))
def test_line_oracle_matches_reference(line_golden, name):
    x, sigma, ref, scaled = line_case(line_golden, name)
    got = O.line_decode_np(x, sigma)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    got4 = O.line_transform_np(x, scale=4, sigma=sigma)
    assert np.array_equal(got4.view(np.uint32), scaled.view(np.uint32))


def test_keypoint_invisible_plane_decodes_to_origin():
    x = np.full((1, 3, 6, 8), -40.0, np.float32)
    out = O.keypoint_decode_np(x, (12, 16))
    assert np.array_equal(out[..., :2], np.zeros((1, 2, 2), np.float32))


def test_line_helpers():
    assert O.calculate_slope_intercept((1.0, 2.0), (1.0, 2.0)) == (None, None)
    k, b = O.calculate_slope_intercept((0.0, 1.0), (2.0, 5.0))
    assert abs(k - 4.0 / (2.0 + 1e-5)) < 1e-12 and abs(b - 1.0) < 1e-12
    assert O.line_eq_intersection((1.0, 0.0), (1.00005, 3.0)) is None          # parallel within 1e-4
    x, y = O.line_eq_intersection((1.0, 0.0), (-1.0, 2.0))
    assert (x, y) == (1.0, 1.0)
    heat = np.zeros((1, 2, 2, 3), np.float32)
    heat[0, 0] = [[10, 20, 0.9], [30, 40, 0.8]]
    heat[0, 1] = [[1, 2, 0.1], [3, 4, 0.05]]
    lines, pts = O.get_line_data(heat, {0: "a", 1: "b"}, scale=4, prob_thre=0.2)
    assert set(lines) == {"a"} and len(pts["a"]) == 2 and pts["b"] == []
