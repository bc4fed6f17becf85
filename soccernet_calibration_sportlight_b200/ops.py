"""Thin torch-tensor wrappers over the C-ABI ops.  PyTorch is only the container for
device memory and the source of the CUDA stream; all arithmetic runs in the
hand-written kernels of csrc/."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib


LAUNCHES = 0          # kernels launched through this module (bench.py reports it)
PROFILE = None        # when a list: (kernel name, start event, end event) per launch


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Launch:
    """Counts one kernel launch and, in profiling mode, brackets it with CUDA events on the
    launching stream."""

    def __init__(self, name: str, device):
        self.name, self.device = name, device

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += 1
        self.ctx = torch.cuda.device(self.device)
        self.ctx.__enter__()
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, self.e0, self.e1))
        self.ctx.__exit__(*exc)
        return False


def _dev(t: torch.Tensor, dtype, what: str) -> int:
    if not t.is_cuda:
        raise _lib.CalError(f"{what}: tensor must live on a CUDA device (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.CalError(f"{what}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.CalError(f"{what}: tensor must be contiguous")
    return t.data_ptr()


def kp_decode(logp: torch.Tensor, size) -> torch.Tensor:
    """(B,C,h,w) fp32 log-probs -> (B,C-1,3) [x,y,conf] (transforms.py:228-239)."""
    B, Cc, h, w = logp.shape
    H, W = int(size[0]), int(size[1])
    if not logp.is_cuda:
        raise _lib.CalError("kp_decode: tensor must live on a CUDA device (no CPU fallback)")
    out = torch.empty((B, max(Cc - 1, 0), 3), dtype=torch.float32, device=logp.device)
    if B == 0 or Cc <= 1:
        return out
    ptr = _dev(logp, torch.float32, "kp_decode")
    with _Launch("kp_decode", logp.device):
        st = _lib.lib().cal_kp_decode(ptr, B, Cc, h, w, H, W,
                                      out.data_ptr(), _stream())
    _lib.check(st, "cal_kp_decode")
    return out


def line_decode(heat: torch.Tensor, sigma: float, scale: float = 1.0) -> torch.Tensor:
    """(B,C,h,w) fp32 probabilities -> (B,C,2,3) two peaks per channel
    (line/transforms.py:224-280)."""
    B, Cc, h, w = heat.shape
    if not heat.is_cuda:
        raise _lib.CalError("line_decode: tensor must live on a CUDA device (no CPU fallback)")
    out = torch.empty((B, Cc, 2, 3), dtype=torch.float32, device=heat.device)
    if B == 0 or Cc == 0:
        return out
    ptr = _dev(heat, torch.float32, "line_decode")
    with _Launch("line_decode", heat.device):
        st = _lib.lib().cal_line_decode(ptr, B, Cc, h, w,
                                        float(sigma), float(scale), out.data_ptr(), _stream())
    _lib.check(st, "cal_line_decode")
    return out


def conv2d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], y: torch.Tensor, *,
           ksize: int, stride: int, cout_rows: int, relu: bool, res: Optional[torch.Tensor] = None,
           mode: int = 0, n_classes: int = 0, cin: int = 0, w_slices: bool = False) -> torch.Tensor:
    """x: fp16 NHWC (B,Hin,Win,Cin_pad); w: fp16 (Cout_rows, taps*Cin_pad);
    y: fp16 NHWC (B,Hout,Wout,Cout_pad) or fp32 NCHW (B,n_classes,Hout,Wout) for mode 1/2."""
    B, Hin, Win, Cin = x.shape
    a = _lib.ConvArgs()
    a.x = _dev(x, torch.float16, "conv2d x")
    a.w = _dev(w, torch.float16, "conv2d w")
    a.bias = _dev(bias, torch.float32, "conv2d bias") if bias is not None else None
    a.res = _dev(res, torch.float16, "conv2d res") if res is not None else None
    if mode == 0:
        a.y = _dev(y, torch.float16, "conv2d y")
        _, Hout, Wout, Cout_pad = y.shape
        if res is not None and tuple(res.shape) != tuple(y.shape):
            raise _lib.CalError("conv2d: residual shape mismatch")
    else:
        a.y = _dev(y, torch.float32, "conv2d y")
        _, ncls, Hout, Wout = y.shape
        if ncls != n_classes:
            raise _lib.CalError("conv2d: n_classes mismatch")
        Cout_pad = 64
    if w_slices:
        if tuple(w.shape) != (ksize * ksize * Cin // 64, cout_rows, 64):
            raise _lib.CalError(f"conv2d: slice-major weight shape {tuple(w.shape)}")
    elif w.shape[0] != cout_rows or w.shape[1] != ksize * ksize * Cin:
        raise _lib.CalError(f"conv2d: weight shape {tuple(w.shape)} vs rows {cout_rows}, K {ksize * ksize * Cin}")
    if bias is not None and bias.numel() != Cout_pad:
        raise _lib.CalError("conv2d: bias length must equal Cout_pad")
    a.B, a.Hin, a.Win, a.Cin_pad = B, Hin, Win, Cin
    a.Hout, a.Wout, a.Cout_pad, a.Cout_rows = Hout, Wout, Cout_pad, cout_rows
    a.ksize, a.stride, a.relu, a.mode, a.n_classes = ksize, stride, int(relu), mode, n_classes
    a.Cin = int(cin)                      # real channels: MMA K steps over the zero pad lanes are skipped
    a.w_slices = int(bool(w_slices))
    name = "conv_tc" if PROFILE is None else f"conv_tc k{ksize}s{stride} {Cin}->{Cout_pad} @{Hout}x{Wout} m{mode}"
    with _Launch(name, x.device):
        st = _lib.lib().cal_conv2d(C.byref(a), _stream())
    _lib.check(st, "cal_conv2d")
    return y


def basicblock(x: torch.Tensor, w1: torch.Tensor, bias1: torch.Tensor, w2: torch.Tensor, bias2: torch.Tensor,
               y: torch.Tensor, *, rows: int, c: int) -> torch.Tensor:
    """BasicBlock.forward (hrnet.py:29-58) in one kernel: x, y fp16 NHWC (B,H,W,64); w1, w2 slice-major
    (9, rows, 64); raises CalError(status CAL_E_UNSUPPORTED) for shapes the two-launch path serves."""
    B, H, W, Cp = x.shape
    if tuple(y.shape) != tuple(x.shape):
        raise _lib.CalError("basicblock: output shape mismatch")
    for w in (w1, w2):
        if tuple(w.shape) != (9, rows, 64):
            raise _lib.CalError(f"basicblock: slice-major weight shape {tuple(w.shape)}")
    a = _lib.BasicBlockArgs()
    a.x = _dev(x, torch.float16, "basicblock x")
    a.w1 = _dev(w1, torch.float16, "basicblock w1")
    a.w2 = _dev(w2, torch.float16, "basicblock w2")
    a.bias1 = _dev(bias1, torch.float32, "basicblock bias1")
    a.bias2 = _dev(bias2, torch.float32, "basicblock bias2")
    a.y = _dev(y, torch.float16, "basicblock y")
    a.B, a.H, a.W, a.C_pad, a.rows, a.C = B, H, W, Cp, rows, int(c)
    name = "basicblock" if PROFILE is None else f"basicblock C{c} @{H}x{W}"
    with _Launch(name, x.device):
        st = _lib.lib().cal_basicblock(C.byref(a), _stream())
    _lib.check(st, "cal_basicblock")
    return y


def stem_conv(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x fp32 NCHW (B,3,H,W) in [0,1], or uint8 NHWC (B,H,W,3) as cv2.imread leaves it (ToTensor's /255 folded
    into the load); w fp32 (64,27); y fp16 NHWC (B,Ho,Wo,64)."""
    u8 = x.dtype == torch.uint8
    if u8:
        B, H, W, _ = x.shape
    else:
        B, _, H, W = x.shape
    _, Ho, Wo, _ = y.shape
    with _Launch("stem_conv", x.device):
        fn = _lib.lib().cal_stem_conv_u8 if u8 else _lib.lib().cal_stem_conv
        st = fn(_dev(x, torch.uint8 if u8 else torch.float32, "stem x"), _dev(w, torch.float32, "stem w"),
                _dev(bias, torch.float32, "stem bias"), _dev(y, torch.float16, "stem y"), B, H, W, Ho, Wo, _stream())
    _lib.check(st, "cal_stem_conv")
    return y


def fuse_combine(y: torch.Tensor, srcs: Sequence[torch.Tensor], bias: Optional[torch.Tensor] = None,
                 relu: bool = False, c: int = 0) -> torch.Tensor:
    """y = [relu](bias + sum_i up_i(src_i)); all fp16 NHWC with the same C_pad; sources of a
    different spatial size are bilinearly resampled (align_corners=True).  c = real channels (0: C_pad): the
    pad lanes are written as zeros without being gathered."""
    B, H, W, Cp = y.shape
    a = _lib.CombineArgs()
    a.y = _dev(y, torch.float16, "combine y")
    a.B, a.H, a.W, a.C_pad = B, H, W, Cp
    a.n_src = len(srcs)
    if not 1 <= len(srcs) <= _lib.CAL_MAX_SOURCES:
        raise _lib.CalError("fuse_combine: 1..6 sources")
    for i, s in enumerate(srcs):
        if s.shape[0] != B or s.shape[3] != Cp:
            raise _lib.CalError("fuse_combine: source batch/channel mismatch")
        a.src[i] = _dev(s, torch.float16, "combine src")
        a.src_h[i], a.src_w[i] = s.shape[1], s.shape[2]
    a.bias = _dev(bias, torch.float32, "combine bias") if bias is not None else None
    a.relu = int(relu)
    a.C = int(c)
    name = "fuse_combine" if PROFILE is None else f"fuse_combine n{len(srcs)} C{Cp} @{H}x{W}"
    with _Launch(name, y.device):
        st = _lib.lib().cal_fuse_combine(C.byref(a), _stream())
    _lib.check(st, "cal_fuse_combine")
    return y


def head_fused(full: torch.Tensor, w_full: torch.Tensor, lows: Sequence[torch.Tensor], bias: torch.Tensor,
               z: Optional[torch.Tensor], cout_rows: int, *, w2: Optional[torch.Tensor] = None,
               bias2: Optional[torch.Tensor] = None, heat: Optional[torch.Tensor] = None,
               mode: int = 0) -> Optional[torch.Tensor]:
    """z = relu(W1_full * full + sum_i up(low_i) + bias) (hrnet.py:489-511 up to the head's ReLU).
    full (B,H,W,64) fp16 NHWC, w_full (rows,64), lows[i] (B,h_i,w_i,Cout_pad), z (B,H,W,Cout_pad).
    With ``w2`` (64, Cout_pad), ``bias2`` (64) and ``heat`` (B, n_classes, H, W) fp32 the final 1x1
    conv + LogSoftmax (mode 1) / Softmax (mode 2) is chained on chip (hrnet.py:325-329) and ``heat``
    is returned instead of z.
    Returns None when the kernel does not support the shape (caller takes the unfused path)."""
    B, H, W, Cf = full.shape
    chain = w2 is not None
    cout_pad = w2.shape[1] if chain else z.shape[3]
    a = _lib.HeadArgs()
    a.full = _dev(full, torch.float16, "head full")
    a.w_full = _dev(w_full, torch.float16, "head w_full")
    if not 1 <= len(lows) <= 4:
        raise _lib.CalError("head_fused: 1..4 low-resolution sources")
    for i, t in enumerate(lows):
        if t.shape[0] != B or t.shape[3] != cout_pad:
            raise _lib.CalError("head_fused: source batch/channel mismatch")
        a.low[i] = _dev(t, torch.float16, "head low")
        a.low_h[i], a.low_w[i] = t.shape[1], t.shape[2]
    a.n_low = len(lows)
    a.bias = _dev(bias, torch.float32, "head bias")
    if chain:
        if tuple(w2.shape) != (64, cout_pad) or bias2.numel() != 64 or tuple(heat.shape[0:1] + heat.shape[2:]) != (B, H, W):
            raise _lib.CalError("head_fused: chained tail shape mismatch")
        a.w2 = _dev(w2, torch.float16, "head w2")
        a.bias2 = _dev(bias2, torch.float32, "head bias2")
        a.heat = _dev(heat, torch.float32, "head heat")
        a.n_classes, a.mode = heat.shape[1], mode
    else:
        a.z = _dev(z, torch.float16, "head z")
        if tuple(z.shape[:3]) != (B, H, W):
            raise _lib.CalError("head_fused: shape mismatch")
    if w_full.shape[1] != Cf or w_full.shape[0] != cout_rows:
        raise _lib.CalError("head_fused: shape mismatch")
    a.B, a.H, a.W, a.Cf_pad, a.Cout_pad, a.Cout_rows = B, H, W, Cf, cout_pad, cout_rows
    name = "head_fused" if PROFILE is None else f"head_fused{'+tail' if chain else ''} {Cf}->{cout_pad} @{H}x{W} n{len(lows)}"
    with _Launch(name, full.device):
        st = _lib.lib().cal_head_fused(C.byref(a), _stream())
    if st == -2:                     # CAL_E_UNSUPPORTED
        global LAUNCHES
        LAUNCHES -= 1
        return None
    _lib.check(st, "cal_head_fused")
    return heat if chain else z


def tma_probe(x: torch.Tensor, box_w: int, box_h: int, estride: int, c0: int, x0: int, y0: int,
              n0: int) -> torch.Tensor:
    """Raw 16 KiB shared-memory image of one activation TMA box (tests only)."""
    B, H, W, Cc = x.shape
    out = torch.empty(16384, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.lib().cal_debug_tma_probe(_dev(x, torch.float16, "probe x"), B, H, W, Cc, box_w, box_h,
                                            estride, c0, x0, y0, n0, out.data_ptr(), _stream())
    _lib.check(st, "cal_debug_tma_probe")
    return out


def camera_solve(preds: torch.Tensor, params: "_lib.SolveParams", line_pts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,57,3) fp32 keypoints [x,y,conf] (+ optional (B,57,2) fp64 line-intersection keypoints,
    NaN = absent) -> (B,16) fp64 camera records, the 128-byte CalCameraRecord viewed as doubles:
    position(3) rotation(9) fx fy rmse, then valid/branch packed as two int32 in the last slot
    (prediction.py:130-136 and the algorithms it dispatches to)."""
    if not preds.is_cuda:
        raise _lib.CalError("camera_solve: tensor must live on a CUDA device (no CPU fallback)")
    B = preds.shape[0]
    if tuple(preds.shape[1:]) != (57, 3):
        raise _lib.CalError(f"camera_solve: preds must be (B,57,3), got {tuple(preds.shape)}")
    out = torch.zeros((B, 16), dtype=torch.float64, device=preds.device)
    if B == 0:
        return out
    lp = None
    if line_pts is not None:
        if tuple(line_pts.shape) != (B, 57, 2):
            raise _lib.CalError("camera_solve: line_pts must be (B,57,2)")
        lp = _dev(line_pts, torch.float64, "camera_solve line_pts")
    with _Launch("camera_solve", preds.device):
        st = _lib.lib().cal_camera_solve(_dev(preds, torch.float32, "camera_solve preds"), lp, C.byref(params), B,
                                         out.data_ptr(), _stream())
    _lib.check(st, "cal_camera_solve")
    return out


def line_points(peaks: torch.Tensor, pair_a: torch.Tensor, pair_b: torch.Tensor, prob_thre: float = 0.0) -> torch.Tensor:
    """(B,23,2,3) decoded line peaks (image pixels) -> (B,57,2) fp64 keypoints from line
    intersections, NaN where absent (export_line_result.py:85-131 + prediction.py:110-124)."""
    B, n_lines = peaks.shape[0], peaks.shape[1]
    out = torch.empty((B, 57, 2), dtype=torch.float64, device=peaks.device)
    if B == 0:
        return out
    with _Launch("line_points", peaks.device):
        st = _lib.lib().cal_line_points(_dev(peaks, torch.float32, "line_points peaks"), B, n_lines,
                                        _dev(pair_a, torch.int32, "pair_a"), _dev(pair_b, torch.int32, "pair_b"),
                                        float(prob_thre), out.data_ptr(), _stream())
    _lib.check(st, "cal_line_points")
    return out


def pnp(obj: torch.Tensor, img: torch.Tensor, K: torch.Tensor, rvec: torch.Tensor, tvec: torch.Tensor,
        refine: bool) -> bool:
    """Single-camera pose: refine=True is Camera.refine_camera's solvePnPRefineLM (camera.py:105-119,
    rvec/tvec updated in place), refine=False is Camera.solve_pnp (camera.py:92-103: what
    cv2.solvePnPRansac returns for these matches - P3P on 4, EPnP on 5, seeded RANSAC + refit on
    more; where OpenCV's RANSAC fails the reference holds uninitialised memory and the least-squares
    pose is returned instead).  All fp64 device tensors: obj (n,3), img (n,2), K (3,3), rvec (3),
    tvec (3).  Returns False only when no finite pose exists."""
    n = obj.shape[0]
    ptrs = [_dev(t, torch.float64, "pnp") for t in (obj, img, K, rvec, tvec)]
    with _Launch("pnp", obj.device):
        if refine:
            st = _lib.lib().cal_pnp_refine(ptrs[0], ptrs[1], n, ptrs[2], ptrs[3], ptrs[4], _stream())
            ok = None
        else:
            ok = torch.zeros(1, dtype=torch.int32, device=obj.device)
            st = _lib.lib().cal_pnp_solve(ptrs[0], ptrs[1], n, ptrs[2], ptrs[3], ptrs[4], ok.data_ptr(), _stream())
    _lib.check(st, "cal_pnp")
    return True if ok is None else int(ok.item()) >= 0
