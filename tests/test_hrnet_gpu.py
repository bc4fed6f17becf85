"""Whole-network parity: the CUDA HRNet (fp16 activations/weights, fp32 accumulation on
tcgen05) against the fp32 CPU oracle (oracle/hrnet_ref.py, pinned bit-exact to the
UNMODIFIED reference modules) on identical seeded weights and frames.

Tolerance (stated here): the reference's predict() is fp32; half-precision
operands perturb the heat maps, so the bar is on the quantities the path consumes:
  * log-probabilities / probabilities: max abs error <= 0.05 in log space (keypoints) and
    <= 5e-3 in probability (lines), relative L2 <= 1e-2;
  * decoded keypoints: integer indices bit-exact wherever the oracle's peak is separated
    from the runner-up by more than the heat-map error bound (all channels of these cases).
The reference's own validation path runs the same network under fp16 autocast
(metamodel.py:67), so half-precision heat maps are in-family for it."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_ref, hrnet_ref as O
from soccernet_calibration_sportlight_b200 import hrnet as P, ops
from tests import inputs as I

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_pair(kind, H, W, B=1, seed=11):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle = O.make_model(kind, seed=seed)
    net = P.HRNetHeatmap(P.w48_config(kind)).load_state_dict(oracle.state_dict()).to(DEV)
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, B, H, W)))
    with torch.no_grad():
        ref = oracle(x)[-1]
    got = net(x.to(DEV))[-1]
    return ref, got.cpu()


@pytest.mark.parametrize("H,W,B", [(96, 160, 2), (135, 241, 1)])
def test_keypoint_net_vs_oracle(H, W, B):
    ref, got = run_pair("keypoints", H, W, B)
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert bool(torch.isfinite(got).all())
    assert float((got - ref).abs().max()) <= 0.05
    assert float((got - ref).norm() / ref.norm()) <= 1e-2
    np.testing.assert_allclose(got.exp().sum(1).numpy(), 1.0, atol=1e-4)


@pytest.mark.parametrize("H,W,B", [(96, 160, 2), (135, 241, 1)])
def test_line_net_vs_oracle(H, W, B):
    ref, got = run_pair("lines", H, W, B)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 5e-3
    assert float((got - ref).norm() / ref.norm()) <= 1e-2
    np.testing.assert_allclose(got.sum(1).numpy(), 1.0, atol=1e-4)


def test_golden_reference_output(golden_dir):
    """Against the stored output of the UNMODIFIED reference module (hrnet_small.npz)."""
    z = np.load(os.path.join(golden_dir, "hrnet_small.npz"))
    for kind, tol in (("keypoints", 0.05), ("lines", 5e-3)):
        oracle = O.make_model(kind, seed=11)
        net = P.HRNetHeatmap(P.w48_config(kind)).load_state_dict(oracle.state_dict()).to(DEV)
        x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, 1, 96, 160))).to(DEV)
        got = net(x)[-1].cpu().numpy()
        assert np.abs(got - z[f"{kind}__out"]).max() <= tol


def test_full_resolution_frame_independence_and_decode():
    """960x540 (BASELINE size): the CPU oracle needs ~4 s/frame, so one frame is compared
    with it and batch entries are checked for frame independence (bit-identical outputs for
    identical frames regardless of batch position)."""
    oracle = O.make_model("keypoints", seed=5)
    net = P.HRNetHeatmap(P.w48_config("keypoints")).load_state_dict(oracle.state_dict()).to(DEV)
    f = I.frames_to_tensor(I.frames_u8(7, 2, 540, 960))
    x = torch.from_numpy(np.stack([f[0], f[1], f[0], f[0]])).to(DEV)
    heat = net(x)[-1]
    assert heat.shape == (4, 58, 270, 480)
    assert torch.equal(heat[0], heat[2]) and torch.equal(heat[0], heat[3])
    with torch.no_grad():
        ref = oracle(torch.from_numpy(f[:1]))[-1]
    assert float((heat[:1].cpu() - ref).abs().max()) <= 0.05
    kp = ops.kp_decode(heat, (540, 960)).cpu().numpy()
    kp_ref = decode_ref.keypoint_decode_np(heat.cpu().numpy(), (540, 960))
    assert np.array_equal(kp.view(np.uint32), kp_ref.view(np.uint32))


def test_load_state_dict_validation():
    net = P.HRNetHeatmap(P.w48_config("lines"))
    sd = dict(P.init_state_dict(P.w48_config("lines"), "lines"))
    bad = dict(sd)
    bad.pop("model.conv1.weight")
    with pytest.raises(KeyError):
        net.load_state_dict(bad)
    bad = dict(sd)
    bad["model.conv1.weight"] = torch.zeros(64, 3, 5, 5)
    with pytest.raises(ValueError):
        net.load_state_dict(bad)
    net.load_state_dict({"_orig_mod." + k if k.startswith("model.conv1") else k: v for k, v in sd.items()})


# ---------------------------------------------------------------------------- end to end
def _peaked_pair(seed_net, seed_frame, seed_cam, contrast):
    """fp32 oracle and CUDA network with the same weights whose heat maps have confident blob peaks
    on one white-noise 960x540 frame (tests/peaked_head.py)."""
    from tests import peaked_head as PH
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle = O.make_model("keypoints", seed=seed_net)
    fr = np.random.default_rng(seed_frame).integers(0, 256, (1, 540, 960, 3), dtype=np.uint8)
    x = torch.from_numpy(I.frames_to_tensor(fr))
    pos, _ = PH.camera_positions(seed_cam)
    PH.calibrate_bn(oracle, x)
    PH.install_templates(oracle, x, pos, contrast=contrast)
    net = P.HRNetHeatmap(P.w48_config("keypoints")).load_state_dict(oracle.state_dict()).to(DEV)
    with torch.no_grad():
        ref = oracle(x)[-1]
    got = net(x.to(DEV))[-1]
    return pos, ref.numpy(), got


@pytest.mark.parametrize("contrast", [8.0, 5.0])
def test_end_to_end_keypoint_index_agreement(contrast):
    """frames -> (57,3) through the fp16-operand CUDA network + CUDA decode against the fp32 oracle
    network + oracle decode, on weights that give confident peaks: how often do the INTEGER keypoint
    indices agree (north_star: bit-exact), and what happens to the camera downstream.
    contrast 8: blobs whose neighbours are ~0.4 log-prob below the peak; contrast 5: flatter blobs (~0.25)
    on a background only just below them."""
    from oracle import camera_ref
    from soccernet_calibration_sportlight_b200 import pitch, prediction
    pos, ref, got = _peaked_pair(5, 7, 3, contrast)
    kp_ref = decode_ref.keypoint_decode_np(ref, (540, 960))[0]
    kp = ops.kp_decode(got, (540, 960)).cpu().numpy()[0]
    err = np.abs(got.cpu().numpy() - ref)[0]                              # (58, 270, 480)
    conf = [c for c in pos if kp_ref[c, 2] > 0.5]
    assert len(conf) >= 12 and all(kp_ref[c, 2] < 1e-6 for c in range(57) if c not in pos)
    agree = [c for c in conf if kp[c, 0] == kp_ref[c, 0] and kp[c, 1] == kp_ref[c, 1]]
    # provably stable peaks: the oracle's best log-prob beats every pixel outside its 3x3 neighbourhood by more
    # than twice the channel's heat-map error
    stable = []
    for c in conf:
        m = ref[0, c]
        r, q = np.unravel_index(np.argmax(m), m.shape)
        rest = m.copy()
        rest[max(0, r - 1):r + 2, max(0, q - 1):q + 2] = -np.inf
        if m[r, q] - rest.max() > 2 * err[c].max():
            stable.append(c)
    shift = max([max(abs(kp[c, 0] - kp_ref[c, 0]), abs(kp[c, 1] - kp_ref[c, 1])) for c in conf] + [0])
    dconf = max(abs(float(kp[c, 2]) - float(kp_ref[c, 2])) for c in conf)
    print(f"\\nend-to-end keypoints (contrast {contrast}): {len(agree)}/{len(conf)} confident keypoints with identical integer "
          f"indices, {len(stable)} provably separated, max index shift {shift:.0f} px, max |conf diff| {dconf:.2e}, "
          f"heat-map max |log-prob err| {err[:57].max():.4f} (confident channels {max(err[c].max() for c in conf):.4f})")
    assert len(agree) >= 0.9 * len(conf)
    assert shift <= 2.0                                                   # never more than one heat-map pixel
    # downstream: the camera from the CUDA path's keypoints against the reference path's
    # (oracle net -> oracle decode -> CameraCreator on cv2)
    kw = {k: v for k, v in camera_ref.MAKE_SUBMIT_KWARGS.items() if k not in ("algorithm", "conf_thresh")}
    creator = prediction.CameraCreator(pitch.PITCH_POINTS, conf_thresh=0.5, algorithm="iterative_voter", **kw)
    cam = creator(kp, "frame")
    rc_creator = camera_ref.make_submit_creator()
    rc = rc_creator(kp_ref, None)
    if rc_creator.pinned and len(agree) == len(conf):
        assert (cam is None) == (rc is None)
        if cam is not None:
            a = np.concatenate([cam.position, cam.rotation.ravel(), [cam.xfocal_length]])
            b = np.concatenate([rc.position, rc.rotation.ravel(), [rc.xfocal_length]])
            rel = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
            print(f"camera from CUDA-path keypoints vs reference path: max relative parameter error {rel:.2e}")
            assert rel < 1e-4


@pytest.mark.parametrize("kind,H,W", [("lines", 540, 960), ("keypoints", 720, 1280), ("lines", 720, 1280),
                                      ("keypoints", 1080, 1920), ("lines", 1080, 1920)])
def test_networks_vs_oracle_at_config5_resolutions(kind, H, W):
    """One frame per network at the BASELINE sizes (540p line net, 720p and 1080p both nets; pyramids end
    in 17 / 23 / 34 rows) against the fp32 oracle."""
    ref, got = run_pair(kind, H, W, 1, seed=13)
    assert got.shape == ref.shape
    tol = 0.05 if kind == "keypoints" else 5e-3
    assert float((got - ref).abs().max()) <= tol
    assert float((got - ref).norm() / ref.norm()) <= 1e-2


@pytest.mark.parametrize("kind", ["w18", "w64", "w48x4"])
def test_other_shipped_configs_vs_oracle(kind, golden_dir):
    """model_config/hrnet_w18.yaml, hrnet_w64.yaml, hrnet_w48x4.yaml through the same engine, against the fp32
    oracle variant (pinned bit-exact to the reference modules) and the stored reference output."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = {"w18": P.w18_config, "w64": P.w64_config, "w48x4": P.w48x4_config}[kind]()
    oracle = O.make_model(kind, seed=11)
    net = P.HRNetHeatmap(cfg).load_state_dict(oracle.state_dict()).to(DEV)
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, 1, 64, 96)))
    got = net(x.to(DEV))[-1].cpu()
    ref = torch.from_numpy(np.load(os.path.join(golden_dir, "hrnet_small.npz"))[f"{kind}__out"])
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 0.05
    x2 = torch.from_numpy(I.frames_to_tensor(I.frames_u8(3, 2, 135, 240)))
    with torch.no_grad():
        ref2 = oracle(x2)[-1]
    got2 = net(x2.to(DEV))[-1].cpu()
    assert float((got2 - ref2).abs().max()) <= 0.05 and float((got2 - ref2).norm() / ref2.norm()) <= 1e-2
