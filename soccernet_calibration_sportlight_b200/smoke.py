"""One small invocation of the hot path on cuda:0, checked against the oracle
(__graft_entry__.smoke)."""
from __future__ import annotations

import numpy as np
import torch


def run() -> None:
    from oracle import decode_ref, hrnet_ref
    from tests import inputs as I
    from . import _lib, hrnet, metamodel, ops

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    _lib.lib()                                         # fails loudly if the .so is missing
    dev = "cuda:0"
    # (1) decode: bit-exact against the oracle
    logp = I.gaussian_logp(4, 2, 58, 20, 24)
    got = ops.kp_decode(torch.from_numpy(logp).to(dev), (40, 48)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), decode_ref.keypoint_decode_np(logp, (40, 48)).view(np.uint32))
    heat = I.two_peak_heat(7, 1, 23, 17, 30)
    gl = ops.line_decode(torch.from_numpy(heat).to(dev), 3.0).cpu().numpy()
    assert np.array_equal(gl.view(np.uint32), decode_ref.line_decode_np(heat, 3.0).view(np.uint32))
    # (2) keypoint network + predict() on a small frame against the fp32 oracle
    oracle = hrnet_ref.make_model("keypoints", seed=3)
    model = metamodel.HRNetMetaModel({"nn_module": {"num_refinement_stages": 0},
                                      "prediction_transform": {"size": (96, 160)}})
    model.nn_module.load_state_dict(oracle.state_dict())
    model.set_device(dev)
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, 1, 96, 160)))
    with torch.no_grad():
        ref = oracle(x)[-1]
    out = model.nn_module(x.to(dev))[-1]
    err = float((out.cpu() - ref).abs().max())
    assert err <= 0.05, f"keypoint heat maps differ from the oracle by {err}"
    pred = model.predict(x)
    assert pred.shape == (1, 57, 3)
    # (3) camera solve on synthetic keypoints against the oracle (reference heuristics on cv2)
    from oracle import camera_ref
    from tests import camera_inputs
    from .pitch import PITCH_POINTS
    from .prediction import CameraCreator
    kps = camera_inputs.clean_predictions(6, seed=11)
    kw = {k: v for k, v in camera_ref.MAKE_SUBMIT_KWARGS.items() if k not in ("algorithm", "conf_thresh")}
    n_cam, worst = 0, 0.0
    for algo in ("opencv_calibration_multiplane", "original_voter"):
        mine = CameraCreator(PITCH_POINTS, conf_thresh=0.5, algorithm=algo, **kw)
        ref = camera_ref.CameraCreatorRef(conf_thresh=0.5, algorithm=algo, **kw)
        cams = mine.batch(kps)
        for i, cam in enumerate(cams):
            rc = ref(kps[i])
            if not ref.pinned:
                continue                               # outcome of the reference itself not reproducible
            assert (cam is None) == (rc is None), f"camera decision differs on frame {i} ({algo})"
            if cam is None:
                continue
            n_cam += 1
            a = np.concatenate([cam.position, cam.rotation.ravel(), [cam.xfocal_length]])
            b = np.concatenate([rc.position, rc.rotation.ravel(), [rc.xfocal_length]])
            worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))))
    assert n_cam > 0 and worst < 1e-4, f"camera parameters differ from the oracle by {worst} (rel)"
    # (4) the camera-solve parity table over the reference's stored outputs (tests/golden/camera_cases.npz)
    from tests import camera_parity

    def gpu_solver(preds, algo, thr):
        t = torch.from_numpy(np.ascontiguousarray(preds, dtype=np.float32)).to(dev)
        rec = ops.camera_solve(t, camera_parity.make_params(algo, thr), None).cpu().numpy()
        flags = rec.view(np.int32).reshape(rec.shape[0], 32)[:, 30:32]
        out = np.zeros((rec.shape[0], 16))
        ok = flags[:, 0] == 1
        out[ok, :14] = rec[ok, :14]
        out[:, 14] = ok
        out[:, 15] = flags[:, 1]
        return out

    stats = camera_parity.compare(gpu_solver)
    print(camera_parity.report(stats))
    camera_parity.assert_parity(stats)
    torch.cuda.synchronize()
    print(f"smoke ok: decode bit-exact, heat-map max|err|={err:.4f}, {n_cam} cameras within {worst:.1e} of the oracle")
