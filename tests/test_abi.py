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
