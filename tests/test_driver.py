"""The pieces either side of the path: checkpoint format, line pickle layouts,
and (GPU) the make_submit-style driver end to end."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from soccernet_calibration_sportlight_b200 import hrnet, make_submit, metamodel, pitch, prediction
from tests import inputs as I


def test_checkpoint_roundtrip_in_the_argus_format(tmp_path):
    """HRNetMetaModel.save writes {'model_name','params','nn_state_dict'} (metamodel.py:88-125);
    load_model reads it back, '_orig_mod.' prefixes (torch.compile) are accepted."""
    params = {"nn_module": {"num_refinement_stages": 0, "hrnet_config": dict(hrnet.w48_config("keypoints"))},
              "prediction_transform": {"size": (540, 960)}}
    m = metamodel.HRNetMetaModel(params)
    sd = hrnet.init_state_dict(hrnet.w48_config("keypoints"), "keypoints", seed=4)
    m.nn_module.load_state_dict(sd)
    path = str(tmp_path / "ckpt.pth")
    m.save(path)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"model_name", "params", "nn_state_dict"} and raw["model_name"] == "HRNetMetaModel"
    assert len(raw["nn_state_dict"]) == 1839                      # w48 keypoint net
    raw["nn_state_dict"] = {"_orig_mod." + k: v for k, v in raw["nn_state_dict"].items()}
    torch.save(raw, path)
    m2 = metamodel.load_model(path)
    for k, v in sd.items():
        assert torch.equal(m2.nn_module.state_dict()[k], v)
    assert m2.prediction_transform.size == (540, 960)
    with pytest.raises(AttributeError):
        m2.predict(torch.zeros(1, 3, 8, 8))                      # no device set: _check_predict_ready


def test_lines_file_both_pickle_layouts(tmp_path):
    """prediction.py:112 reads entry['lines'][0]; export_line_result.py:188 stores the dict itself."""
    lines = {"Side line left": (50.0, -100.0), "Side line top": (0.01, 30.0), "Big rect. left top": (0.02, 120.0)}
    kw = {k: v for k, v in prediction.MAKE_SUBMIT_KWARGS.items()}
    for layout in ("dict", "list"):
        path = str(tmp_path / f"lines_{layout}.pkl")
        entry = {"lines": lines if layout == "dict" else [lines], "points": {}}
        with open(path, "wb") as f:
            pickle.dump({"img_0.jpg": entry}, f)
        cc = prediction.CameraCreator(pitch.PITCH_POINTS, lines_file=path, **kw)
        pts = cc._get_points_from_lines("img_0.jpg")
        assert set(pts) == {11, 13}                               # (side line left x big rect top), (x side line top)
        x, y = prediction.line_eq_intersection(lines["Side line left"], lines["Side line top"])
        assert pts[13] == (x, y)
        assert cc._get_points_from_lines("other.jpg") == {}
    with pytest.raises(AssertionError):
        prediction.CameraCreator(pitch.PITCH_POINTS, lines_file=str(tmp_path / "missing.pkl"), **kw)


def test_frames_to_tensor_is_totensor():
    fr = I.frames_u8(3, 2, 12, 16)
    t = make_submit.frames_to_tensor(list(fr))
    assert t.shape == (2, 3, 12, 16) and t.dtype == torch.float32
    assert np.array_equal(t.numpy(), I.frames_to_tensor(fr))


@pytest.mark.gpu
def test_make_submit_driver_end_to_end(tmp_path):
    """jpg frames -> checkpoint -> predict -> batched camera solve -> camera_*.json in the
    reference's schema (camera.py:162-174).  Random weights give no confident keypoint, so the
    calibrator is fed through a model stub whose predict() returns synthetic keypoints - the
    network itself is exercised once to check the plumbing."""
    import cv2
    from tests import camera_inputs as CI
    img_dir, save_dir = tmp_path / "imgs", tmp_path / "out"
    img_dir.mkdir()
    frames = I.frames_u8(5, 3, 96, 160)
    for i in range(3):
        cv2.imwrite(str(img_dir / f"{i:05d}.jpg"), frames[i])
    params = {"nn_module": {"num_refinement_stages": 0}, "prediction_transform": {"size": (540, 960)}}
    model = metamodel.HRNetMetaModel(params).set_device("cuda:0")
    ckpt = str(tmp_path / "kp.pth")
    model.save(ckpt)
    comp = make_submit.main(["--model", ckpt, "--img-dir", str(img_dir), "--save-dir", str(save_dir), "--batch-size", "2"])
    assert comp == 0.0 and os.listdir(save_dir) == []            # conf ~ 1/58: nothing to calibrate

    kps = CI.clean_predictions(3, seed=1)

    class Stub:
        def predict(self, x):
            n = x.shape[0]
            out, self.i = torch.from_numpy(kps[self.i:self.i + n]).cuda(), self.i + n
            return out
    stub = Stub()
    stub.i = 0
    cal = prediction.CameraCreator(pitch.PITCH_POINTS, **prediction.MAKE_SUBMIT_KWARGS)
    comp = make_submit.run(stub, cal, str(img_dir), str(save_dir), batch_size=2, quiet=True)
    files = sorted(os.listdir(save_dir))
    assert comp > 0 and files and all(f.startswith("camera_") and f.endswith(".json") for f in files)
    d = json.load(open(save_dir / files[0]))
    assert set(d) == {"pan_degrees", "tilt_degrees", "roll_degrees", "position_meters", "x_focal_length",
                      "y_focal_length", "principal_point", "radial_distortion", "tangential_distortion",
                      "thin_prism_distortion"}
    assert d["principal_point"] == [480.0, 270.0] and len(d["position_meters"]) == 3


@pytest.mark.gpu
def test_export_line_result_round_trip(tmp_path):
    """jpg frames -> line checkpoint -> export_line_result (export_line_result.py:134-201) -> pickle ->
    CameraCreator(lines_file=...) (prediction.py:104-124): the writer's layout is what the reader takes,
    and its lines equal the oracle's get_line_data on the same decoded peaks."""
    import pickle
    import cv2
    from oracle import decode_ref
    from soccernet_calibration_sportlight_b200 import export_line_result
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    frames = I.frames_u8(9, 3, 96, 160)
    for i in range(3):
        cv2.imwrite(str(img_dir / f"{i:05d}.jpg"), frames[i])
    model = metamodel.LineMetaModel({"nn_module": {"num_refinement_stages": 0},
                                     "prediction_transform": {"scale": 4, "sigma": 3.0}}).set_device("cuda:0")
    ckpt, out = str(tmp_path / "line.pth"), str(tmp_path / "res" / "lines.pkl")
    model.save(ckpt)
    res = export_line_result.main(["--model", ckpt, "--image-folder", str(img_dir), "--result-file", out,
                                   "--prob-thre", "0.0", "--batch-size", "2", "--no-vis"])
    stored = pickle.load(open(out, "rb"))
    assert sorted(stored) == ["00000.jpg", "00001.jpg", "00002.jpg"] and set(stored["00000.jpg"]) == {"lines", "points"}
    # against the oracle's dictionary building on the decoded peaks of the same (jpg-decoded) frame
    img = cv2.imread(str(img_dir / "00001.jpg"), cv2.IMREAD_COLOR)
    heat = model.nn_module(make_submit.frames_to_tensor([img]).cuda())[-1]
    peaks = decode_ref.line_decode_np(heat.cpu().numpy(), 3.0)
    lines, points = decode_ref.get_line_data(peaks, pitch.LINE_CLS, scale=4, prob_thre=0.0)
    assert set(lines) == set(stored["00001.jpg"]["lines"]) and len(lines) == 23
    for k, (slope, icpt) in lines.items():
        s2, i2 = stored["00001.jpg"]["lines"][k]
        assert (slope is None) == (s2 is None)
        if slope is not None:
            assert float(slope) == float(s2) and float(icpt) == float(i2)
    assert res["00001.jpg"]["points"].keys() == points.keys()
    # the reader takes the writer's layout.  A class whose two peaks coincide is stored as (None, None)
    # (export_line_result.py:69-70) and makes CameraCreator.__init__ raise, here as in the reference
    # (prediction.py:116-119): with random weights that can happen, so drop such entries first
    clean = {n: {"lines": {k: v for k, v in e["lines"].items() if v[0] is not None}, "points": e["points"]}
             for n, e in stored.items()}
    had_none = any(v[0] is None for e in stored.values() for v in e["lines"].values())
    out2 = str(tmp_path / "res" / "lines_clean.pkl")
    export_line_result.write(clean, out2)
    cal = prediction.CameraCreator(pitch.PITCH_POINTS, lines_file=out2, **prediction.MAKE_SUBMIT_KWARGS)
    assert isinstance(cal.lines_data, dict)
    if had_none:
        with pytest.raises(TypeError):
            prediction.CameraCreator(pitch.PITCH_POINTS, lines_file=out, **prediction.MAKE_SUBMIT_KWARGS)
