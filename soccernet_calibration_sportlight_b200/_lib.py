"""ctypes binding of the C-ABI library (include/calib_b200.h).  There is no CPU
fallback: if the shared library is missing the import fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcalib_b200.so")

CAL_MAX_SOURCES = 6

EXPORTS = [
    "cal_abi_version", "cal_last_error", "cal_set_smem_headroom", "cal_kp_decode", "cal_line_decode", "cal_conv2d", "cal_basicblock",
    "cal_stem_conv", "cal_stem_conv_u8", "cal_fuse_combine", "cal_head_fused", "cal_camera_solve", "cal_pnp_refine", "cal_pnp_solve",
    "cal_line_points", "cal_evaluate_cameras", "cal_hrnet_weight_count", "cal_hrnet_create", "cal_hrnet_forward",
    "cal_hrnet_output_shape", "cal_hrnet_launches", "cal_hrnet_destroy",
    "cal_debug_tma_probe", "cal_debug_shift_mma", "cal_debug_mn_mma", "cal_debug_mma_rate", "cal_debug_mma_pattern",
]


class CalError(RuntimeError):
    pass


class ConvArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("res", C.c_void_p),
                ("y", C.c_void_p),
                ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Cin_pad", C.c_int32),
                ("Hout", C.c_int32), ("Wout", C.c_int32), ("Cout_pad", C.c_int32),
                ("Cout_rows", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
                ("relu", C.c_int32), ("mode", C.c_int32), ("n_classes", C.c_int32), ("Cin", C.c_int32), ("w_slices", C.c_int32)]


class BasicBlockArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("w1", C.c_void_p), ("bias1", C.c_void_p), ("w2", C.c_void_p), ("bias2", C.c_void_p),
                ("y", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C_pad", C.c_int32), ("rows", C.c_int32), ("C", C.c_int32)]


class CombineArgs(C.Structure):
    _fields_ = [("y", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C_pad", C.c_int32),
                ("n_src", C.c_int32),
                ("src", C.c_void_p * CAL_MAX_SOURCES),
                ("src_h", C.c_int32 * CAL_MAX_SOURCES),
                ("src_w", C.c_int32 * CAL_MAX_SOURCES),
                ("bias", C.c_void_p), ("relu", C.c_int32), ("C", C.c_int32)]


class HeadArgs(C.Structure):
    _fields_ = [("full", C.c_void_p), ("w_full", C.c_void_p), ("low", C.c_void_p * 4),
                ("low_h", C.c_int32 * 4), ("low_w", C.c_int32 * 4), ("n_low", C.c_int32),
                ("bias", C.c_void_p), ("z", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cf_pad", C.c_int32),
                ("Cout_pad", C.c_int32), ("Cout_rows", C.c_int32),
                ("w2", C.c_void_p), ("bias2", C.c_void_p), ("heat", C.c_void_p),
                ("n_classes", C.c_int32), ("mode", C.c_int32)]


class HrnetStage(C.Structure):
    _fields_ = [("num_modules", C.c_int32), ("num_branches", C.c_int32), ("block_type", C.c_int32),
                ("num_blocks", C.c_int32 * 4), ("num_channels", C.c_int32 * 4)]


class HrnetConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("num_classes", C.c_int32), ("stem_width", C.c_int32), ("upscale", C.c_int32),
                ("stage", HrnetStage * 4)]


class SolveParams(C.Structure):
    _fields_ = [("pitch_xyz", C.c_double * (57 * 3)),
                ("algorithm", C.c_int32), ("img_w", C.c_int32), ("img_h", C.c_int32),
                ("conf_thresh", C.c_float), ("conf_threshs", C.c_float * 8),
                ("n_conf_threshs", C.c_int32),
                ("min_points", C.c_int32), ("min_points_per_plane", C.c_int32),
                ("min_points_for_refinement", C.c_int32), ("reliable_thresh", C.c_int32),
                ("min_focal_length", C.c_float), ("max_rmse", C.c_float), ("max_rmse_rel", C.c_float)]


class CameraRecord(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("rotation", C.c_double * 9),
                ("fx", C.c_double), ("fy", C.c_double), ("rmse", C.c_double),
                ("valid", C.c_int32), ("branch", C.c_int32)]


assert C.sizeof(CameraRecord) == 128

_lib = None


def lib() -> C.CDLL:
    """Load (once) and type the library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CalError(f"{LIB_PATH} not built: run `python __graft_entry__.py` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, f64 = C.c_void_p, C.c_int, C.c_float, C.c_double
    L.cal_abi_version.restype = C.c_int
    L.cal_last_error.restype = C.c_char_p
    L.cal_kp_decode.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp]
    L.cal_line_decode.argtypes = [vp, i32, i32, i32, i32, f64, f32, vp, vp]
    L.cal_conv2d.argtypes = [C.POINTER(ConvArgs), vp]
    L.cal_basicblock.argtypes = [C.POINTER(BasicBlockArgs), vp]
    L.cal_stem_conv.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.cal_stem_conv_u8.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.cal_fuse_combine.argtypes = [C.POINTER(CombineArgs), vp]
    L.cal_head_fused.argtypes = [C.POINTER(HeadArgs), vp]
    L.cal_debug_tma_probe.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    L.cal_debug_shift_mma.argtypes = [vp, vp, i32, i32, vp, vp]
    L.cal_debug_mn_mma.argtypes = [vp, vp, i32, vp, vp]
    L.cal_debug_mma_rate.argtypes = [i32, i32, i32, i32, vp, vp]
    L.cal_debug_mma_pattern.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp]
    L.cal_camera_solve.argtypes = [vp, vp, C.POINTER(SolveParams), i32, vp, vp]
    L.cal_pnp_refine.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.cal_pnp_solve.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.cal_line_points.argtypes = [vp, i32, i32, vp, vp, f32, vp, vp]
    L.cal_hrnet_weight_count.argtypes = [C.POINTER(HrnetConfig), C.POINTER(C.c_size_t)]
    L.cal_hrnet_create.argtypes = [C.POINTER(HrnetConfig), vp, C.c_size_t, C.POINTER(vp)]
    L.cal_hrnet_forward.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.cal_hrnet_output_shape.argtypes = [vp, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.cal_hrnet_launches.argtypes = [vp]
    L.cal_hrnet_destroy.argtypes = [vp]
    L.cal_evaluate_cameras.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, f64, vp, vp, i32, i32, vp, vp, vp]
    for name in EXPORTS:
        if hasattr(L, name) and name != "cal_last_error":
            getattr(L, name).restype = C.c_int
    L.cal_hrnet_launches.restype = C.c_long
    _lib = L
    return L


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().cal_last_error().decode("utf-8", "replace")
        raise CalError(f"{what} failed with status {status}: {msg}")
