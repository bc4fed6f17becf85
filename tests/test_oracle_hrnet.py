"""oracle/hrnet_ref.py against outputs of the UNMODIFIED reference networks
(tests/golden/make_golden_hrnet.py loaded one state_dict into both and required
bit-identical outputs in the build container; here the stored reference output pins the
oracle on any machine) and the state_dict schema of the product's architecture walk."""
import json
import os

import numpy as np
import torch

from oracle import hrnet_ref as O
from soccernet_calibration_sportlight_b200 import hrnet as P
from tests import inputs as I


def test_oracle_reproduces_reference_outputs(golden_dir):
    z = np.load(os.path.join(golden_dir, "hrnet_small.npz"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, 1, 96, 160)))
    for kind in ("keypoints", "lines"):
        m = O.make_model(kind, seed=11)
        with torch.no_grad():
            y = m(x)[-1].numpy()
        ref = z[f"{kind}__out"]
        assert y.shape == ref.shape
        # same torch build -> bit-identical; other builds may reorder fp32 sums
        np.testing.assert_allclose(y, ref, rtol=0, atol=2e-4)
        if kind == "keypoints":
            np.testing.assert_allclose(np.exp(y).sum(1), 1.0, atol=1e-4)
        else:
            np.testing.assert_allclose(y.sum(1), 1.0, atol=1e-4)


def test_state_dict_schema_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "hrnet_state_keys.json")) as f:
        keys = json.load(f)
    for kind in ("keypoints", "lines"):
        schema = P.state_dict_schema(P.w48_config(kind), kind)
        assert list(schema) == list(keys[kind])
        assert all(list(schema[k]) == keys[kind][k] for k in schema)
        oracle_sd = O.make_model(kind, seed=1).state_dict()
        assert list(oracle_sd) == list(schema)


def test_conv_inventory_matches_survey():
    """307 convs / 507.82 GFLOP per 960x540 frame for the keypoint net."""
    for kind, n_convs, gflop in (("keypoints", 307, 507.82), ("lines", 307, 371.4)):
        net = P._walk(P.w48_config(kind), kind)
        convs = P._all_convs(net)
        assert len(convs) == n_convs
        assert abs(P.conv_gflop_per_frame(kind, 540, 960) - gflop) < 0.05


def test_config_access():
    cfg = P.w48_config("keypoints")
    assert "upscale" in cfg and cfg.upscale == 2 and cfg.stage4.num_channels[-1] == 384
    assert "upscale" not in P.w48_config("lines")


def test_other_shipped_configs(golden_dir):
    """hrnet_w18 / hrnet_w64 / hrnet_w48x4 (model_config/*.yaml): the oracle's variants reproduce the stored
    reference outputs, and the product's architecture walk yields the reference's state_dict schema."""
    z = np.load(os.path.join(golden_dir, "hrnet_small.npz"))
    with open(os.path.join(golden_dir, "hrnet_state_keys.json")) as f:
        keys = json.load(f)
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, 1, 64, 96)))
    for kind, cfg in (("w18", P.w18_config()), ("w64", P.w64_config()), ("w48x4", P.w48x4_config())):
        schema = P.state_dict_schema(cfg, "keypoints")
        assert list(schema) == list(keys[kind]) and all(list(schema[k]) == keys[kind][k] for k in schema)
        if kind == "w64":
            continue                               # 117 M parameters: the forward is left to the GPU suite
        with torch.no_grad():
            y = O.make_model(kind, seed=11)(x)[-1].numpy()
        np.testing.assert_allclose(y, z[f"{kind}__out"], rtol=0, atol=2e-4)
    assert P.w48x4_config().upscale == 4 and P.w18_config().stage1.num_channels == [32]
