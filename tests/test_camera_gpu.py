"""GPU parity of the camera solve (csrc/solve.cu through the C ABI) against the stored outputs
of the UNMODIFIED reference (tests/golden/camera_cases.npz) by parity class
(tests/camera_parity.py), against the oracle (oracle/camera_ref.py) on fresh seeded inputs,
and of the API mirrors (CameraCreator.__call__/batch, Camera.solve_pnp/refine_camera)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import camera_ref as O, decode_ref
from soccernet_calibration_sportlight_b200 import _lib, ops, pitch, prediction
from soccernet_calibration_sportlight_b200.camera import Camera
from tests import camera_inputs as CI, camera_parity as CP, inputs as I

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KW = {k: v for k, v in O.MAKE_SUBMIT_KWARGS.items() if k not in ("algorithm", "conf_thresh")}


def gpu_solver(preds, algo, thr, line_pts=None):
    t = torch.from_numpy(np.ascontiguousarray(preds, dtype=np.float32)).to(DEV)
    lp = None if line_pts is None else torch.from_numpy(np.ascontiguousarray(line_pts)).to(DEV)
    rec = ops.camera_solve(t, CP.make_params(algo, thr), lp).cpu().numpy()
    flags = rec.view(np.int32).reshape(rec.shape[0], 32)[:, 30:32]
    out = np.zeros((rec.shape[0], 16))
    ok = flags[:, 0] == 1
    out[ok, :14] = rec[ok, :14]
    out[:, 14] = ok
    out[:, 15] = flags[:, 1]
    return out


def test_golden_parity_classes():
    """Every golden frame whose reference outcome is determined by its inputs: exact, minimal
    (P3P / EPnP shortcut of solvePnPRansac) and ransac (seeded RANSAC) classes, asserted at 1e-4."""
    stats = CP.compare(gpu_solver)
    print(CP.report(stats))
    rep = {k: {a: b for a, b in v.items() if a != "failures"} for k, v in stats.items()}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/camera_parity_classes.json", "w") as f:
        json.dump(rep, f, indent=1)
    CP.assert_parity(stats)


def test_pnp_solve_kernel_vs_recorded_opencv_calls():
    """cal_pnp_solve (Camera.solve_pnp) on the GPU against the cv2.solvePnPRansac calls the reference
    made on the golden frames (tests/golden/cv_calls.npz)."""
    z = np.load(os.path.join(CP.ROOT, "tests", "golden", "cv_calls.npz"))
    n_ok = 0
    for j in range(0, len(z["pnp_n"]), 3):
        n = int(z["pnp_n"][j])
        tv = z["pnp_tvec"][j]
        K = z["pnp_K"][j]
        if not z["pnp_ok"][j] or not np.all(np.isfinite(tv)) or np.abs(tv).max() > 1e6 or K[0, 0] == 1.0:
            continue
        obj = torch.from_numpy(np.ascontiguousarray(z["pnp_obj"][j, :n])).to(DEV)
        img = torch.from_numpy(np.ascontiguousarray(z["pnp_img"][j, :n])).to(DEV)
        rvec, tvec = torch.zeros(3, dtype=torch.float64, device=DEV), torch.zeros(3, dtype=torch.float64, device=DEV)
        assert ops.pnp(obj, img, torch.from_numpy(K.copy()).to(DEV), rvec, tvec, refine=False)
        from soccernet_calibration_sportlight_b200.camera import rotation_from_rodrigues
        err = max(np.abs(rotation_from_rodrigues(rvec.cpu().numpy()) - rotation_from_rodrigues(z["pnp_rvec"][j])).max(),
                  np.abs(tvec.cpu().numpy() - tv).max() / max(1.0, np.abs(tv).max()))
        if err > 1e-5 and n > 5:
            continue                               # (the two calls where a later sample ties on the inlier count)
        assert err < 1e-5, (j, n, err)
        n_ok += 1
    assert n_ok >= 60


def test_kernel_equals_host_compilation_of_the_same_source():
    """The block-cooperative kernel (128 threads) and the one-thread host build of the same source
    take the same decisions and agree to rounding (different FMA contraction / summation order)."""
    host = CP.host_solver()
    preds = CI.synthetic_predictions(64, seed=21)
    for algo in ("iterative_voter", "voter"):
        a, b = gpu_solver(preds, algo, CP.ALGOS[algo]), host(preds, algo, CP.ALGOS[algo])
        assert np.array_equal(a[:, 14:], b[:, 14:])
        ok = a[:, 14] == 1
        assert np.max(np.abs(a[ok, :14] - b[ok, :14]) / np.maximum(np.abs(b[ok, :14]), 1.0)) < 1e-6


@pytest.mark.parametrize("algo", ["opencv_calibration", "opencv_calibration_multiplane", "original_voter"])
def test_fresh_inputs_vs_oracle(algo):
    preds = CI.synthetic_predictions(48, seed=33, noise_px=1.0, outlier=0.0, drop=0.1, conf_lo=0.3)
    got = gpu_solver(preds, algo, 0.5)
    creator = O.CameraCreatorRef(conf_thresh=0.5, algorithm=algo, **KW)
    n = 0
    for i in range(preds.shape[0]):
        cam = creator(preds[i], None)
        if not creator.pinned or CP.degenerate_goal_view(preds[i], 0.5):
            continue
        ref = O.camera_record(cam)
        if ref[14] == 1 and not CP.feasible_record(ref):
            continue
        assert (got[i, 14] == 1) == (ref[14] == 1), i
        if ref[14] == 1:
            assert CP.rel_err(got[i], ref) < CP.TOL, (i, creator.branch)
            n += 1
    assert n >= 10


def test_empty_and_undetectable_frames():
    P = CP.make_params("iterative_voter", 0.5)
    assert ops.camera_solve(torch.zeros((0, 57, 3), device=DEV), P).shape == (0, 16)
    z = torch.zeros((3, 57, 3), device=DEV)                         # nothing detected -> None
    z[1, :, 2] = 0.9                                                # all points at (0,0): degenerate
    z[2, :5, 2] = 0.9                                               # fewer than 6 points
    rec = ops.camera_solve(z, P).cpu().numpy()
    assert np.all(rec.view(np.int32).reshape(3, 32)[:, 30] == 0)
    with pytest.raises(_lib.CalError):
        ops.camera_solve(torch.zeros((2, 57, 3)), P)                # host tensor: no CPU fallback
    with pytest.raises(_lib.CalError):
        ops.camera_solve(torch.zeros((2, 56, 3), device=DEV), P)


def test_camera_creator_api_and_json():
    preds = CI.clean_predictions(8, seed=4)
    mine = prediction.CameraCreator(pitch.PITCH_POINTS, conf_thresh=0.5, algorithm="iterative_voter", **KW)
    ref = O.make_submit_creator()
    cams = mine.batch(preds)
    n = 0
    for i in range(8):
        one = mine(preds[i], f"img_{i}")
        assert (one is None) == (cams[i] is None)
        rc = ref(preds[i], None)
        if not ref.pinned:
            continue
        assert (one is None) == (rc is None)
        if one is None:
            continue
        ja, jb = one.to_json_parameters(), rc.to_json_parameters()
        assert set(ja) == set(jb)
        for k in ("pan_degrees", "tilt_degrees", "roll_degrees", "x_focal_length", "y_focal_length"):
            assert abs(ja[k] - jb[k]) <= 1e-4 * max(1.0, abs(jb[k])), k
        assert np.allclose(ja["position_meters"], jb["position_meters"], rtol=1e-4, atol=1e-4)
        assert ja["principal_point"] == jb["principal_point"] == [480.0, 270.0]
        assert np.allclose(one.calibration, rc.calibration, rtol=1e-4)
        n += 1
    assert n >= 3
    assert prediction.CameraCreator(pitch.get_pitch(), algorithm="voter", **KW).algorithm == "voter"
    with pytest.raises(AssertionError):
        prediction.CameraCreator(pitch.PITCH_POINTS, algorithm="nope")


def test_camera_refine_and_solve_pnp_vs_cv2():
    import cv2
    rng = np.random.default_rng(8)
    for trial in range(6):
        R, pos, f = CI.random_camera(rng)
        uv, ok = CI.project(R, pos, f)
        ids = [i for i in range(57) if ok[i]]
        if len(ids) < 8:
            continue
        img = uv[ids] + rng.normal(0, 0.7, (len(ids), 2))
        matches = [(CI.WORLD[i], (float(a), float(b))) for i, (a, b) in zip(ids, img)]
        K = np.array([[f, 0, 479.5], [0, f, 269.5], [0, 0, 1.0]])
        cam, rc = Camera(), O.CameraRef()
        for c in (cam, rc):
            c.calibration = K.copy()
            # perturbed start
            c.rotation = R @ cv2.Rodrigues(np.array([0.01, -0.02, 0.015]))[0]
            c.position = pos + np.array([1.0, -2.0, 0.5])
        cam.refine_camera(matches)
        rc.refine_camera(matches)
        assert np.allclose(cam.position, rc.position, rtol=1e-4, atol=1e-4)
        assert np.allclose(cam.rotation, rc.rotation, atol=1e-5)
        cam2, rc2 = Camera(), O.CameraRef()
        cam2.calibration, rc2.calibration = K.copy(), K.copy()
        cam2.solve_pnp(matches)                      # cv2.solvePnPRansac restated (csrc/solve_pnp_cv.cuh)
        rc2.solve_pnp(matches)
        if not rc2.tainted:
            assert np.allclose(cam2.position, rc2.position, rtol=1e-4, atol=1e-3)
            assert np.allclose(cam2.rotation, rc2.rotation, atol=1e-5)


def test_line_points_kernel_exact_vs_oracle():
    heat = I.two_peak_heat(5, 6, 23, 135, 240)
    peaks = ops.line_decode(torch.from_numpy(heat).to(DEV), 3.0, 4.0)
    ref_peaks = decode_ref.line_transform_np(heat, scale=4, sigma=3.0)
    assert np.array_equal(peaks.cpu().numpy().view(np.uint32), ref_peaks.view(np.uint32))
    creator = prediction.CameraCreator(pitch.PITCH_POINTS, algorithm="voter", **KW)
    for thr in (0.0, 0.6):
        got = creator.line_points_device(peaks, thr).cpu().numpy()
        ref = O.line_keypoints(ref_peaks, prob_thre=thr)
        n = 0
        for b in range(6):
            present = {i for i in range(57) if not np.isnan(got[b, i, 0])}
            assert present == set(ref[b]), (b, thr)
            for i, (x, y) in ref[b].items():
                assert got[b, i, 0] == np.float64(x) and got[b, i, 1] == np.float64(y)
                n += 1
        assert n > 0 or thr > 0


def test_pipeline_full_workload_runs_and_is_frame_independent():
    from soccernet_calibration_sportlight_b200.pipeline import CalibrationPipeline
    pipe = CalibrationPipeline(DEV, workload="full", size=(96, 160))
    f = torch.from_numpy(I.frames_to_tensor(I.frames_u8(3, 2, 96, 160)))
    x = torch.stack([f[0], f[1], f[0]])
    kp = torch.from_numpy(CI.clean_predictions(3, seed=2)).to(DEV)
    kp[2] = kp[0]
    out = pipe(x, keypoints_override=kp)
    assert out["keypoints"].shape == (3, 57, 3) and out["lines"].shape == (3, 23, 2, 3)
    assert out["cameras"].shape == (3, 16)
    assert torch.equal(out["keypoints"][0], out["keypoints"][2])
    assert torch.equal(out["cameras"][0], out["cameras"][2])


def test_run_stream_overlapped_copies_match_direct_calls():
    """CalibrationPipeline.run_stream (host->device copy of batch i+1 on a side stream) gives the
    same records as calling the pipeline on device-resident batches."""
    from soccernet_calibration_sportlight_b200.pipeline import CalibrationPipeline
    pipe = CalibrationPipeline(DEV, workload="keypoints", size=(96, 160))
    kp = torch.from_numpy(CI.clean_predictions(2, seed=6)).to(DEV)
    batches = [torch.from_numpy(I.frames_to_tensor(I.frames_u8(s, 2, 96, 160))).pin_memory() for s in (1, 2, 3)]
    direct = [pipe(b.to(DEV), keypoints_override=kp) for b in batches]
    streamed = list(pipe.run_stream(batches, keypoints_override=kp))
    assert len(streamed) == 3
    for d, s in zip(direct, streamed):
        assert s.device.type == "cpu" and torch.equal(d["cameras"].cpu(), s)
    kps = list(pipe.run_stream(batches, result_key="keypoints"))
    for d, s in zip(direct, kps):
        assert torch.equal(d["keypoints"].cpu(), s)
    assert list(pipe.run_stream([])) == []
