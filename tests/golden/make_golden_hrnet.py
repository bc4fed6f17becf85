"""Pins oracle/hrnet_ref.py to the UNMODIFIED reference networks (build container only):
    python tests/golden/make_golden_hrnet.py
Reference modules run here: src/models/hrnet/model.py:130-150 (HRNetHeatmap, w48, 58 maps)
and src/models/line/model.py (HRNetHeatmap, w48, 23 maps).  One seeded state_dict is loaded
into both the reference module and the oracle restatement; CPU outputs must be bit-identical.
Writes hrnet_state_keys.json (state_dict schema) and hrnet_small.npz (reference outputs on
a small frame, for the oracle self-check on other machines)."""
import json
import os
import sys

import numpy as np
import torch

import refimport

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
refimport.setup()
from src.models.hrnet.model import HRNetHeatmap as RefKp  # noqa: E402
from src.models.line.model import HRNetHeatmap as RefLine  # noqa: E402

from oracle import hrnet_ref as O  # noqa: E402
from soccernet_calibration_sportlight_b200 import hrnet as P  # noqa: E402
from tests import inputs as I  # noqa: E402

torch.set_num_threads(8)
keys = {}
out = {}
for kind, Ref in (("keypoints", RefKp), ("lines", RefLine)):
    cfg = P.w48_config(kind)
    ref = Ref(cfg, 0, cfg.num_classes).eval()
    oracle = O.make_model(kind, seed=11)
    sd = oracle.state_dict()
    ref.load_state_dict(sd, strict=True)
    keys[kind] = {k: list(v.shape) for k, v in ref.state_dict().items()}
    schema = P.state_dict_schema(cfg, kind)
    assert list(schema.keys()) == list(ref.state_dict().keys()), "schema key order differs"
    assert all(tuple(ref.state_dict()[k].shape) == tuple(s) for k, s in schema.items())
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, 1, 96, 160)))
    with torch.no_grad():
        yr = ref(x)[-1]
        yo = oracle(x)[-1]
    assert torch.equal(yr, yo), f"{kind}: oracle != reference (max {float((yr - yo).abs().max())})"
    print(kind, "oracle == reference bit-exact on", tuple(x.shape), "->", tuple(yr.shape),
          "params", sum(p.numel() for p in ref.parameters()))
    out[f"{kind}__out"] = yr.numpy()
# the other shipped keypoint configs (model_config/hrnet_w18.yaml, hrnet_w64.yaml, hrnet_w48x4.yaml), read
# from the reference's own yaml files: oracle == reference bit for bit; a small output slice is stored
CFG_DIR = os.path.join(refimport.REF, "src", "models", "hrnet", "model_config")
for kind, fn in (("w18", "hrnet_w18.yaml"), ("w64", "hrnet_w64.yaml"), ("w48x4", "hrnet_w48x4.yaml")):
    cfg = P.config_from_yaml(os.path.join(CFG_DIR, fn))
    ref = RefKp(cfg, 0, cfg.num_classes).eval()
    oracle = O.make_model(kind, seed=11)
    ref.load_state_dict(oracle.state_dict(), strict=True)
    schema = P.state_dict_schema(cfg, "keypoints")
    assert list(schema.keys()) == list(ref.state_dict().keys()), f"{kind}: schema key order differs"
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, 1, 64, 96)))
    with torch.no_grad():
        yr = ref(x)[-1]
        yo = oracle(x)[-1]
    assert torch.equal(yr, yo), f"{kind}: oracle != reference"
    print(kind, "oracle == reference bit-exact ->", tuple(yr.shape), "params", sum(p.numel() for p in ref.parameters()))
    out[f"{kind}__out"] = yr.numpy()
    keys[kind] = {k: list(v.shape) for k, v in ref.state_dict().items()}
with open(os.path.join(HERE, "hrnet_state_keys.json"), "w") as f:
    json.dump(keys, f)
np.savez_compressed(os.path.join(HERE, "hrnet_small.npz"), **out)
print("written")
