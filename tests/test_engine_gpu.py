"""The library's own engine (csrc/engine.cu behind cal_hrnet_create / cal_hrnet_forward) against its readable
twin, the op-by-op schedule in hrnet.py: same kernels in the same order, so the outputs must be bit-identical -
for both networks, odd sizes, uint8 frames and the other shipped configurations."""
import numpy as np
import pytest
import torch

from soccernet_calibration_sportlight_b200 import hrnet as P, ops
from tests import inputs as I

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def pair(cfg, kind, seed=3):
    a = P.HRNetHeatmap(cfg, kind=kind)
    a.load_state_dict(P.init_state_dict(cfg, kind, seed=seed))
    b = P.HRNetHeatmap(cfg, kind=kind).load_state_dict(a.state_dict())
    a.use_engine, b.use_engine = True, False
    return a.to(DEV), b.to(DEV)


@pytest.mark.parametrize("kind,H,W,B", [("keypoints", 96, 160, 2), ("lines", 96, 160, 2), ("keypoints", 135, 241, 1),
                                        ("lines", 135, 241, 1), ("keypoints", 270, 480, 3)])
def test_engine_equals_python_schedule(kind, H, W, B):
    eng, py = pair(P.w48_config(kind), kind)
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(5, B, H, W))).to(DEV)
    n0 = ops.LAUNCHES
    a = eng(x)[-1]
    n_eng = ops.LAUNCHES - n0
    n0 = ops.LAUNCHES
    b = py(x)[-1]
    n_py = ops.LAUNCHES - n0
    assert a.shape == b.shape and torch.equal(a, b)
    assert n_eng == n_py and n_eng > 250                     # the same launches, counted on both sides
    assert torch.equal(eng(x)[-1], a)                        # and again (allocator reuse)


def test_engine_uint8_frames():
    eng, py = pair(P.w48_config("keypoints"), "keypoints")
    fr = I.frames_u8(9, 2, 96, 160)
    xf = torch.from_numpy(I.frames_to_tensor(fr)).to(DEV)
    x8 = torch.from_numpy(fr).to(DEV)
    a, b, c = eng(x8)[-1], eng(xf)[-1], py(x8)[-1]
    assert torch.equal(a, b) and torch.equal(a, c)
    with pytest.raises(ValueError):
        eng(torch.zeros((2, 3, 96, 160), dtype=torch.uint8, device=DEV))


@pytest.mark.parametrize("name", ["w18", "w64", "w48x4"])
def test_engine_other_configs(name):
    cfg = {"w18": P.w18_config, "w64": P.w64_config, "w48x4": P.w48x4_config}[name]()
    eng, py = pair(cfg, "keypoints")
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(2, 1, 64, 96))).to(DEV)
    assert torch.equal(eng(x)[-1], py(x)[-1])
