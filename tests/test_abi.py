"""The C-ABI library loads on a CPU-only box and exports every entry point declared in
include/calib_b200.h; argument validation that needs no device works; there is no CPU
fallback (compute without CUDA returns CAL_E_CUDA or raises)."""
import ctypes
import os
import re

import pytest

from soccernet_calibration_sportlight_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "calib_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cal_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 10
    L = ctypes.CDLL(built_lib)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in include/calib_b200.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == syms


def test_abi_version_and_struct_sizes(built_lib):
    L = _lib.lib()
    assert L.cal_abi_version() == 1
    assert ctypes.sizeof(_lib.CameraRecord) == 128


def test_argument_validation_without_device(built_lib):
    L = _lib.lib()
    assert L.cal_conv2d(None, None) == -1
    assert b"null" in L.cal_last_error()
    assert L.cal_kp_decode(None, 1, 0, 4, 4, 8, 8, None, None) == -1       # C < 1
    assert L.cal_kp_decode(None, 0, 58, 4, 4, 8, 8, None, None) == 0       # empty batch is a no-op
    assert L.cal_line_decode(None, 0, 23, 4, 4, 3.0, 1.0, None, None) == 0


def test_no_cpu_fallback():
    import torch
    from soccernet_calibration_sportlight_b200 import hrnet, ops
    with pytest.raises(_lib.CalError):
        ops.kp_decode(torch.zeros(1, 3, 4, 4), (8, 8))
    with pytest.raises(RuntimeError):
        hrnet.HRNetHeatmap(hrnet.w48_config("keypoints")).to("cpu")


def _w48(kind: int):
    """CalHrnetConfig of model_config/hrnet_w48.yaml, written out by hand (no hrnet.py)."""
    c = _lib.HrnetConfig()
    c.kind, c.num_classes, c.stem_width, c.upscale = kind, (58 if kind == 0 else 23), 64, (2 if kind == 0 else 1)
    spec = [(1, 1, 1, [4], [64]), (1, 2, 0, [4, 4], [48, 96]), (4, 3, 0, [4, 4, 4], [48, 96, 192]),
            (3, 4, 0, [4, 4, 4, 4], [48, 96, 192, 384])]
    for i, (nm, nb, bt, blocks, ch) in enumerate(spec):
        c.stage[i].num_modules, c.stage[i].num_branches, c.stage[i].block_type = nm, nb, bt
        for b in range(nb):
            c.stage[i].num_blocks[b], c.stage[i].num_channels[b] = blocks[b], ch[b]
    return c


def test_engine_walk_counts_the_reference_state_dict(built_lib):
    """cal_hrnet_weight_count (the library's own architecture walk, host only) against the parameter and
    BatchNorm-statistics count of the reference's state_dict (tests/golden/hrnet_state_keys.json)."""
    import json
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "hrnet_state_keys.json")))
    L = _lib.lib()
    for kind, name in ((0, "keypoints"), (1, "lines")):
        want = sum(int(__import__("math").prod(s)) for k, s in keys[name].items() if not k.endswith("num_batches_tracked"))
        n = ctypes.c_size_t()
        cfg = _w48(kind)
        assert L.cal_hrnet_weight_count(ctypes.byref(cfg), ctypes.byref(n)) == 0
        assert n.value == want
    bad = _w48(0)
    bad.stage[1].num_branches = 3
    h = ctypes.c_void_p()
    assert L.cal_hrnet_create(ctypes.byref(bad), ctypes.c_void_p(1), 0, ctypes.byref(h)) == -1


@pytest.mark.gpu
def test_forward_through_the_c_abi_alone():
    """A network forward driven by ctypes only - cal_hrnet_create / cal_hrnet_forward / cal_hrnet_destroy, the
    configuration written by hand, the weights straight from a state_dict - against the fp32 oracle."""
    import numpy as np
    import torch
    from oracle import hrnet_ref
    from tests import inputs as I
    L = _lib.lib()
    for kind, name, tol in ((0, "keypoints", 0.05), (1, "lines", 5e-3)):
        oracle = hrnet_ref.make_model(name, seed=11)
        blob = torch.cat([v.reshape(-1).float() for k, v in oracle.state_dict().items()
                          if not k.endswith("num_batches_tracked")]).contiguous()
        cfg, h = _w48(kind), ctypes.c_void_p()
        assert L.cal_hrnet_create(ctypes.byref(cfg), blob.data_ptr(), blob.numel(), ctypes.byref(h)) == 0, L.cal_last_error()
        frames = I.frames_u8(21, 2, 96, 160)
        x = torch.from_numpy(I.frames_to_tensor(frames))
        nc, oh, ow = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        assert L.cal_hrnet_output_shape(h, 96, 160, ctypes.byref(nc), ctypes.byref(oh), ctypes.byref(ow)) == 0
        heat = torch.empty((2, nc.value, oh.value, ow.value), device="cuda:0")
        xd = x.contiguous().cuda()
        st = torch.cuda.current_stream().cuda_stream
        assert L.cal_hrnet_forward(h, xd.data_ptr(), 0, 2, 96, 160, heat.data_ptr(), st) == 0, L.cal_last_error()
        with torch.no_grad():
            ref = oracle(x)[-1]
        assert heat.shape == ref.shape and float((heat.cpu() - ref).abs().max()) <= tol
        # uint8 HWC frames (cv2.imread layout): bit-identical to the float route
        heat8 = torch.empty_like(heat)
        x8 = torch.from_numpy(frames).cuda()
        assert L.cal_hrnet_forward(h, x8.data_ptr(), 1, 2, 96, 160, heat8.data_ptr(), st) == 0, L.cal_last_error()
        assert torch.equal(heat8, heat)
        assert L.cal_hrnet_launches(h) > 500
        assert L.cal_hrnet_destroy(h) == 0
    assert L.cal_hrnet_forward(None, None, 0, 1, 96, 160, None, None) == -1
