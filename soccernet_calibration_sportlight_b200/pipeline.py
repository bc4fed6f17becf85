"""The calibration hot path as one object: frames -> keypoints (-> line peaks) -> cameras.

Mirrors the reference's inference driver loop body (src/utils/make_submit.py:59-73):
``model.predict(tensor)`` followed by ``CameraCreator`` on every frame - here the camera
solve is the batched CUDA kernel (``CameraCreator.batch``) instead of a 16-process CPU pool,
and the optional line network (src/utils/export_line_result.py:176-185) runs in the same
pass so its intersections can feed the solve (prediction.py:104-124) without a pickle
round-trip.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import metamodel
from .hrnet import w48_config


class CalibrationPipeline:
    """workload 'kp_decode': keypoint net + decode only (BASELINE config 2);
    'keypoints': keypoint net + decode + camera solve (what make_submit.py runs,
    LINES_FILE=None); 'full': + line net + two-peak decode feeding line intersections.

    ``size`` = (H, W) of the image the keypoints, line peaks and cameras are expressed in (the
    reference's prediction-transform ``size`` and CameraCreator ``img_size``, 540 x 960 everywhere in
    its configs); the network input may have any resolution of the same aspect ratio (BASELINE
    config 5): both decoders rescale to ``size``."""

    def __init__(self, device="cuda:0", workload: str = "keypoints", size=(540, 960),
                 kp_state_dict=None, line_state_dict=None, camera_kwargs: Optional[dict] = None,
                 line_sigma: float = 3.0, seed: int = 0):
        if workload not in ("kp_decode", "keypoints", "full"):
            raise ValueError(workload)
        self.device = torch.device(device)
        self.two_streams = os.environ.get("CAL_TWO_STREAMS", "1") != "0"
        # the camera solve of batch i under the networks of batch i+1 (the reference overlaps them too:
        # a 16-process CPU pool solves while the GPU runs on, make_submit.py:53-73).  The solve kernel is
        # 64 small blocks (64 threads x <= 168 registers, 39 KB) that slip into the gaps the networks leave:
        # kernel tails, the CUDA-core kernels (stem, multi-resolution sums), and any SM whose persistent
        # conv CTA leaves room.  `solve_headroom` (CAL_SMEM_HEADROOM) makes the conv kernels keep that many
        # bytes of shared memory free on every SM; measured on B200 the gaps alone hide the whole solve
        # (93.6 ms per step with 0, 94.0 ms with 40960, 129 ms without the overlap), so the default is 0.
        self.overlap_solve = os.environ.get("CAL_SOLVE_OVERLAP", "1") != "0"
        self.solve_headroom = int(os.environ.get("CAL_SMEM_HEADROOM", "0"))
        from . import _lib
        _lib.check(_lib.lib().cal_set_smem_headroom(self.solve_headroom), "cal_set_smem_headroom")
        self._side = None
        self._solve_stream = None
        self._copy = None
        self.workload = workload
        self.size = (int(size[0]), int(size[1]))
        self.kp_model = metamodel.HRNetMetaModel({"nn_module": {"num_refinement_stages": 0},
                                                  "prediction_transform": {"size": self.size}})
        if kp_state_dict is not None:
            self.kp_model.nn_module.load_state_dict(kp_state_dict)
        self.kp_model.set_device(self.device)
        self.line_model = None
        if workload == "full":
            self.line_model = metamodel.LineMetaModel({"nn_module": {"num_refinement_stages": 0},
                                                       "prediction_transform": {"scale": 4, "sigma": line_sigma}})
            if line_state_dict is not None:
                self.line_model.nn_module.load_state_dict(line_state_dict)
            self.line_model.set_device(self.device)
        self.camera_creator = None
        if workload != "kp_decode":
            from .prediction import CameraCreator, MAKE_SUBMIT_KWARGS
            from .pitch import PITCH_POINTS
            kw = dict(MAKE_SUBMIT_KWARGS if camera_kwargs is None else camera_kwargs)
            kw.setdefault("img_size", (self.size[1], self.size[0]))       # (W, H), prediction.py:44
            self.camera_creator = CameraCreator(PITCH_POINTS, **kw)

    @torch.no_grad()
    def __call__(self, frames: torch.Tensor, keypoints_override: Optional[torch.Tensor] = None,
                 defer_solve: bool = False) -> Dict[str, torch.Tensor]:
        """frames: (B,3,H,W) fp32 in [0,1] BGR (the reference's tensor), or (B,H,W,3) uint8 BGR as cv2.imread leaves
        them (ToTensor's /255 then happens in the stem kernel: a quarter of the bytes to move), on the host
        (pinned) or on the device.
        Returns device tensors: 'keypoints' (B,57,3), optionally 'lines' (B,23,2,3), and
        'cameras' (B,16) fp64 records (see prediction.CameraCreator.batch_records).
        ``keypoints_override`` (B,57,3) feeds the camera solve instead of the network's own
        keypoints (benchmarks with random-init weights, whose confidences never pass a threshold).
        ``defer_solve``: the solve is enqueued on the pipeline's solve stream and the current stream does
        NOT wait for it - the next call's networks run over it; out['solve_stream'] is that stream (work
        that consumes 'cameras' goes there, or waits for out['cameras_ready'])."""
        x = frames.to(self.device, non_blocking=True)
        line_pts = None
        if self.line_model is not None:
            # line heat maps are a quarter of the network input: peaks -> `size` coordinates
            in_h, in_w = (x.shape[1], x.shape[2]) if x.dtype == torch.uint8 else (x.shape[-2], x.shape[-1])
            sy, sx = self.size[0] / in_h, self.size[1] / in_w
            if abs(sx - sy) > 1e-3 * sx:
                raise ValueError(f"network input {(in_h, in_w)} and size {self.size} differ in aspect ratio")
            self.line_model.prediction_transform.scale = 4.0 * sx
        if self.line_model is not None and self.two_streams:
            # the two networks are independent: on two streams the ramp-up / tail of every
            # (persistent, one CTA per SM) kernel of one network is filled by the other's
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                lines = self.line_model.predict(x)
            out = {"keypoints": self.kp_model.predict(x)}
            main.wait_stream(self._side)
            x.record_stream(self._side)
            lines.record_stream(main)
            out["lines"] = lines
            line_pts = self.camera_creator.line_points_device(out["lines"])
        else:
            out = {"keypoints": self.kp_model.predict(x)}
            if self.line_model is not None:
                out["lines"] = self.line_model.predict(x)
                line_pts = self.camera_creator.line_points_device(out["lines"])
        if self.camera_creator is not None:
            kp = out["keypoints"] if keypoints_override is None else keypoints_override
            if not self.overlap_solve:
                out["cameras"] = self.camera_creator.batch_records(kp, line_pts)
                return out
            main = torch.cuda.current_stream(self.device)
            if self._solve_stream is None:
                self._solve_stream = torch.cuda.Stream(device=self.device)
            side = self._solve_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out["cameras"] = self.camera_creator.batch_records(kp, line_pts)
                ready = torch.cuda.Event()
                ready.record(side)
            kp.record_stream(side)
            if line_pts is not None:
                line_pts.record_stream(side)
            out["solve_stream"], out["cameras_ready"] = side, ready
            if not defer_solve:
                main.wait_stream(side)
                out["cameras"].record_stream(main)
        return out

    @torch.no_grad()
    def run_stream(self, host_batches, keypoints_override: Optional[torch.Tensor] = None, result_key: Optional[str] = None,
                   to_host: bool = True):
        """Generator over an iterable of pinned HOST frame batches; yields each batch's result
        (``result_key`` defaults to 'cameras' when the solve is on, else 'keypoints') on the host.
        This is the loop ``make_submit.py:59-73`` runs, software-pipelined so that nothing waits on the
        critical path: the host->device copy of batch i+1 runs on a copy stream while batch i computes,
        the camera solve of batch i runs under the networks of batch i+1, and batch i's result travels to
        a pinned host buffer asynchronously - it is yielded one iteration later, when the host has
        already enqueued batch i+1.  With ``to_host=False`` device tensors are yielded (the caller
        synchronises with the solve stream: the tensor carries ``.cal_ready``, a CUDA event)."""
        key = result_key or ("cameras" if self.camera_creator is not None else "keypoints")
        if self._copy is None:                      # one copy stream per pipeline (torch hands streams out of a small pool)
            self._copy = torch.cuda.Stream(device=self.device)
        copy_stream = self._copy
        compute = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        # two device staging buffers, allocated once (no allocator traffic inside the loop): buffer k
        # is refilled for batch i+2 only after batch i's kernels have consumed it
        bufs, consumed = [None, None], [None, None]

        def stage(h, k):
            with torch.cuda.stream(copy_stream):
                if bufs[k] is None or bufs[k].shape != h.shape or bufs[k].dtype != h.dtype:
                    bufs[k] = torch.empty(h.shape, dtype=h.dtype, device=self.device)
                    bufs[k].record_stream(compute)
                    consumed[k] = None
                if consumed[k] is not None:
                    copy_stream.wait_event(consumed[k])
                bufs[k].copy_(h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return bufs[k], ev, k
        try:
            nxt = stage(next(it), 0)
        except StopIteration:
            return
        pending = None                              # (pinned host tensor or device tensor, event) of the previous batch
        host_out = [None, None]
        n = 0
        while nxt is not None:
            cur, ev, k = nxt
            try:
                nxt = stage(next(it), 1 - k)
            except StopIteration:
                nxt = None
            compute.wait_event(ev)
            out = self(cur, keypoints_override=keypoints_override, defer_solve=True)
            consumed[k] = torch.cuda.Event()
            consumed[k].record(compute)
            res = out[key]
            src_stream = out.get("solve_stream", compute) if key == "cameras" else compute
            if to_host:
                slot = n & 1
                if host_out[slot] is None or host_out[slot].shape != res.shape or host_out[slot].dtype != res.dtype:
                    host_out[slot] = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
                with torch.cuda.stream(src_stream):
                    host_out[slot].copy_(res, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(src_stream)
                res.record_stream(src_stream)
                item = (host_out[slot], done)
            else:
                done = out.get("cameras_ready") if key == "cameras" else None
                if done is None:
                    done = torch.cuda.Event()
                    done.record(compute)
                item = (res, done)
            n += 1
            if pending is not None:
                pending[1].synchronize()
                yield pending[0].clone() if to_host else pending[0]
            pending = item
        if pending is not None:
            pending[1].synchronize()
            yield pending[0].clone() if to_host else pending[0]
