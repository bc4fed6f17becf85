#!/usr/bin/env python
"""Benchmark of the calibration hot path (BASELINE.json metric: calibrated frames/s at 960x540).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic 960x540 frames per GPU.
Under torchrun (N > 1) every rank processes its own shard of the batch (weak scaling) and
one NCCL all-gather collects the per-frame results.  Rank 0 prints ONE JSON line.

  value    : frames/s with the frames already resident in HBM (CUDA events, max over ranks)
  e2e      : the same through the public API from pinned HOST frames, host->device and
             device->host copies inside the timed region
  roofline : the dominant kernel (tcgen05 implicit-GEMM conv) - algorithmic FLOPs of the
             reference's convolutions / summed device time of its launches (per-launch
             CUDA events on the launching stream, one extra profiled step) against the
             measured bf16/fp16 tensor peak of MEASURED_PEAKS.json; the decode kernel's
             HBM figure is reported beside it under "decode"
  cpu_baseline : the oracle (CPU restatement of the reference) timed on this box's host
             cores on a bounded sample (rank 0, N = 1 only)

--impl reference times the reference's own CPU path (the oracle port: the Python reference
cannot travel to the GPU box) on a bounded sample per step, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name -> (description, networks, camera solve)
    "kp_decode": ("batch=64 HRNet-w48 keypoint forward + heatmap decode, 960x540 (BASELINE config 2)", ("keypoints",), False),
    "keypoints": ("batch=64 HRNet-w48 keypoints + decode + camera solve, 960x540 (make_submit.py path)", ("keypoints",), True),
    "full": ("batch=64 full pipeline: keypoint net + line net + decodes + camera solve, 960x540 (BASELINE config 3)",
             ("keypoints", "lines"), True),
}
H_IMG, W_IMG = 540, 960
BATCH_PER_GPU = 64
METRIC = "calibrated frames/sec @960x540"
UNIT = "frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor_burst=float(d["bf16_tflops"]),
                    tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 200 ms while active."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------- the reference CPU path
class CpuPath:
    """The oracle (CPU restatement of the reference path, oracle/): networks in fp32 through
    torch CPU kernels, decodes, camera solve through cv2 - what prediction.py + camera.py +
    the reference modules do on the host."""

    def __init__(self, workload: str, threads: int):
        import torch
        from oracle import decode_ref, hrnet_ref
        torch.set_num_threads(threads)
        self.torch, self.decode_ref, self.hrnet_ref = torch, decode_ref, hrnet_ref
        _, nets, solve = WORKLOADS[workload]
        self.kp = hrnet_ref.make_model("keypoints", seed=0)
        self.ln = hrnet_ref.make_model("lines", seed=1) if "lines" in nets else None
        self.creator = None
        if solve:
            from oracle import camera_ref
            self.camera_ref = camera_ref
            self.creator = camera_ref.make_submit_creator()

    def __call__(self, x, synth_preds=None):
        torch = self.torch
        with torch.no_grad():
            preds = self.hrnet_ref.predict(self.kp, x, (H_IMG, W_IMG)).numpy()
            line_kp = None
            if self.ln is not None:
                heat = self.ln(x)[-1].numpy()
                line_kp = self.camera_ref.line_keypoints(self.decode_ref.line_transform_np(heat, scale=4, sigma=3.0))
        cams = None
        if self.creator is not None:
            # random weights give conf ~ 1/58 < every threshold, so the
            # solve is timed on synthetic keypoints of the same shape
            src = preds if synth_preds is None else synth_preds
            cams = [self.camera_ref.solve(self.creator, src[i], None if line_kp is None else line_kp[i])
                    for i in range(src.shape[0])]
        return preds, cams


def run_reference(args, rank):
    """--impl reference: the oracle port on the host cores, one frame per step."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from tests import inputs as I
    cores = os.cpu_count() or 1
    path = CpuPath(args.workload, cores)
    frames = torch.from_numpy(I.frames_to_tensor(I.frames_u8(1, 1, *args.net_hw)))
    synth = synthetic_keypoints(args.steps + args.warmup, seed=5) if WORKLOADS[args.workload][2] else None
    for i in range(args.warmup):
        path(frames, None if synth is None else synth[i:i + 1])
    t0 = time.perf_counter()
    for i in range(args.steps):
        path(frames, None if synth is None else synth[args.warmup + i:args.warmup + i + 1])
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"1 frame per step x {args.steps} steps through the oracle port (torch CPU fp32 + numpy + cv2)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][0], "name": args.workload, "frames_per_step": 1,
                   "note": "one CPU process on rank 0's host cores whatever --gpus is: the ratio to an N-GPU line "
                           "compares N GPUs with one host"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def synthetic_keypoints(n, seed=0):
    """(n,57,3) keypoint predictions of plausible broadcast cameras (tests/inputs.py)."""
    from tests import camera_inputs
    return camera_inputs.synthetic_predictions(n, seed=seed)


# ------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from soccernet_calibration_sportlight_b200 import _lib, hrnet as P, ops
    from soccernet_calibration_sportlight_b200.pipeline import CalibrationPipeline

    _lib.lib()                                   # fail loudly if the CUDA library is missing
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    desc, nets, solve = WORKLOADS[args.workload]
    B = args.batch
    pipe = CalibrationPipeline(dev, workload=args.workload, size=(H_IMG, W_IMG))

    g = torch.Generator().manual_seed(1234 + rank)
    nh, nw = args.net_hw
    # frames as cv2.imread leaves them (make_submit.py:66): uint8 HWC BGR; ToTensor's /255 runs in the stem kernel
    # (--float-frames: the reference's float CHW tensor instead, four times the bytes)
    u8 = torch.randint(0, 256, (B, nh, nw, 3), generator=g, dtype=torch.uint8)
    host = (u8.permute(0, 3, 1, 2).float().div_(255.0).contiguous() if args.float_frames else u8).pin_memory()
    frames = host.to(dev)
    synth = None
    if solve:
        # random-init weights give conf ~ 1/58 everywhere, i.e. every frame would exit the solve at
        # the first gate; the solve therefore runs on synthetic keypoints of plausible cameras
        # (the network's decoded keypoints are still produced and gathered)
        synth = torch.from_numpy(synthetic_keypoints(B, seed=100 + rank)).to(dev)

    def result_of(out):
        return out["cameras"] if "cameras" in out else out["keypoints"]

    from soccernet_calibration_sportlight_b200 import sharding

    def step(x):
        # the solve of this batch is enqueued on the pipeline's solve stream and runs under the next
        # batch's networks (pipeline.overlap_solve); the timed region ends with a device-wide synchronize
        out = pipe(x, keypoints_override=synth, defer_solve=True) if solve else pipe(x)
        res = result_of(out)
        if world > 1:                             # the one collective of the path: gather the records
            st = out.get("solve_stream")
            if st is not None:
                with torch.cuda.stream(st):
                    return sharding.all_gather_records(res, world * B)
            return sharding.all_gather_records(res, world * B)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between two events on the compute stream; the second event waits for the solve stream too, so the
        camera solve of the LAST batch - which has no next batch to hide under - is inside the timed region.
        Returns (ms including that drain, ms up to the last network kernel)."""
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if pipe._solve_stream is not None:
            torch.cuda.current_stream().wait_stream(pipe._solve_stream)
        e2.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0].item()), float(ms[1].item())

    for _ in range(max(args.warmup, 3)):
        step(frames)
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    n0 = ops.LAUNCHES
    ms, ms_nets = timed(lambda: step(frames), args.steps)
    launches = ops.LAUNCHES - n0
    clocks = sampler.stop() if sampler else None

    # end to end through the public API: pinned host frames -> device, pipeline, result -> host,
    # every step (CalibrationPipeline.run_stream overlaps the copy of step i+1 with step i; with
    # several ranks the gather of the records closes each step)
    res_host = None

    def e2e_run(steps):
        nonlocal res_host
        for r in pipe.run_stream((host for _ in range(steps)), keypoints_override=synth, to_host=(world == 1)):
            if world > 1:
                r = sharding.all_gather_records(r, world * B).to("cpu")
            res_host = r
    e2e_run(2)
    # the box's pinned host -> device rate on this batch (explains e2e when a box's PCIe path is slow:
    # the copy of step i+1 overlaps step i only while it is shorter than the step)
    hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    hb0.record()
    _tmp = host.to(dev, non_blocking=True)
    hb1.record()
    torch.cuda.synchronize()
    h2d_ms = hb0.elapsed_time(hb1)
    del _tmp
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item())
    h2d = host.numel() * host.element_size()
    d2h = res_host.numel() * res_host.element_size()

    # N > 1, outside every timed region: the gathered keypoints and camera records must equal, bit for bit,
    # what rank 0 computes ALONE for the same frames (every rank's inputs are regenerated from their seeds)
    mg_check = None
    if world > 1:
        out = pipe(frames, keypoints_override=synth) if solve else pipe(frames)
        mine = [out["keypoints"].contiguous()] + ([out["cameras"].contiguous()] if solve else [])
        gathered = []
        for t in mine:
            g_all = torch.empty((world * B,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            dist.all_gather_into_tensor(g_all, t)
            gathered.append(g_all)
        if rank == 0:
            same = [True] * len(mine)
            for r in range(world):
                gr = torch.Generator().manual_seed(1234 + r)
                fr = torch.randint(0, 256, (B, nh, nw, 3), generator=gr, dtype=torch.uint8)
                fr = (fr.permute(0, 3, 1, 2).float().div_(255.0).contiguous() if args.float_frames else fr).to(dev)
                sy = torch.from_numpy(synthetic_keypoints(B, seed=100 + r)).to(dev) if solve else None
                o = pipe(fr, keypoints_override=sy) if solve else pipe(fr)
                alone = [o["keypoints"]] + ([o["cameras"]] if solve else [])
                for k, (a, g_all) in enumerate(zip(alone, gathered)):
                    same[k] = same[k] and bool(torch.equal(a.contiguous().view(torch.uint8), g_all[r * B:(r + 1) * B].view(torch.uint8)))
            mg_check = {"ranks": world, "what": "NCCL-gathered results of all ranks vs rank 0 computing every rank's frames alone",
                        "keypoints_bit_identical": same[0], "records_bit_identical": same[1] if solve else None}
        barrier()

    # per-kernel device time: one extra step with CUDA events around every launch, the two networks
    # on ONE stream so that a launch's event pair times that launch alone (in the timed region the
    # networks run on two streams and their kernels overlap)
    two_streams, pipe.two_streams = pipe.two_streams, False
    step(frames)
    torch.cuda.synchronize()
    ops.PROFILE = []
    step(frames)
    torch.cuda.synchronize()
    pipe.two_streams = two_streams
    per, by_shape = {}, {}
    for name, a, b in ops.PROFILE:
        t = a.elapsed_time(b)
        d = per.setdefault(name.split(" ")[0], [0.0, 0])
        d[0] += t
        d[1] += 1
        d = by_shape.setdefault(name, [0.0, 0])
        d[0] += t
        d[1] += 1
    ops.PROFILE = None
    if rank == 0 and args.shapes_out:
        with open(args.shapes_out, "w") as f:
            f.write("kernel and shape,launches,total_ms,avg_us\n")
            for k, (t, n) in sorted(by_shape.items(), key=lambda kv: -kv[1][0]):
                f.write(f"{k},{n},{t:.3f},{t / n * 1e3:.1f}\n")
    pk = peaks()
    # every tcgen05 conv launch: generic + halo kernels ("conv_tc") and the fused head
    conv_keys = [k for k in per if k in ("conv_tc", "basicblock") or k.startswith("head_fused")]   # "head_fused", "head_fused+tail"
    conv_ms = sum(per[k][0] for k in conv_keys)
    conv_n = sum(per[k][1] for k in conv_keys)
    gflop_frame = sum(P.conv_gflop_per_frame(k, nh, nw) for k in nets)
    stem_gflop = 2.0 * 3 * 64 * 9 * ((nh + 1) // 2) * ((nw + 1) // 2) / 1e9 * len(nets)   # conv1 runs on CUDA cores (stem_conv)
    conv_tflop = (gflop_frame - stem_gflop) * B / 1e3
    achieved = conv_tflop / (conv_ms / 1e3) if conv_ms > 0 else 0.0
    roof = {"kernel": "tcgen05 implicit-GEMM convs (basicblock_kernel + conv3x3_pair_kernel + conv3x3_halo_kernel + conv_tc_kernel + head_chain_kernel / head_fused_kernel, all "
                      "launches of one step)", "bound": "tensor",
            "achieved": achieved, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["tensor_sustained"], "peak_source": pk["source"] + " (sustained fp16/bf16 dense)",
            "launches_per_step": conv_n, "avg_launch_ms": conv_ms / max(conv_n, 1),
            "algorithmic_tflop_per_step": conv_tflop, "share_of_step": conv_ms / (ms / args.steps), "traffic": None,
            "timing": "CUDA events around every launch of one extra single-stream step (the timed steps overlap the "
                      "two networks on two streams, so share_of_step can exceed what a serial step would show)"}
    # counter-backed DRAM traffic of the step's largest conv bucket (one ncu --set full capture, tools/ncu_shapes.py ->
    # tools/summarize_profiles.py -> profiles/traffic.json), per launch, next to the algorithmic bytes of that launch
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and (nh, nw) == (H_IMG, W_IMG) and B == 64:
        tj = json.load(open(tpath))
        if tj.get("launches"):
            top = tj["launches"][0]
            roof["traffic"] = top["dram_bytes"]
            roof["traffic_detail"] = {"launch": f"{top['kernel']} {top['shape']} (the largest bucket of the step)",
                                      "algorithmic_bytes": top["algorithmic_bytes"], "source": tj["source"],
                                      "all": [{k: l[k] for k in ("shape", "dram_bytes", "algorithmic_bytes")} for l in tj["launches"]]}
    dec = {}
    if "kp_decode" in per:
        bytes_alg = B * 57 * ((nh + 1) // 2) * ((nw + 1) // 2) * 4
        t = per["kp_decode"][0] / 1e3
        dec = {"kernel": "kp_decode_vec_kernel", "bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": pk["hbm"],
               "unit": "GB/s", "frac": bytes_alg / t / 1e9 / pk["hbm"], "ms": per["kp_decode"][0]}
    kernels_ms = {k: round(v[0], 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])}

    if rank != 0:
        return
    frames_total = world * B * args.steps
    line = {
        "metric": METRIC, "value": frames_total / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "ms_per_step_steady_state": ms_nets / args.steps,      # without the last batch's un-hidden camera solve (the limit for long runs)
        "vs_baseline": None, "dtype": "f16 operands, f32 accumulate (tcgen05); f32 decode; f64 camera solve",
        "data": "synthetic",
        "config": {"workload": desc if (nh, nw) == (H_IMG, W_IMG) else
                   desc.replace("960x540", f"{nw}x{nh} network input, keypoints in 960x540 coordinates").replace(
                       "BASELINE config 2", "BASELINE config 5 sweep").replace("BASELINE config 3", "BASELINE config 5 sweep"),
                   "name": args.workload, "batch_per_gpu": B, "global_batch": world * B,
                   "resolution": [nw, nh], "weights": "random-init HRNet-w48 (seeded)",
                   "frames": ("float32 CHW in [0,1] (the reference's ToTensor output)" if args.float_frames else
                              "uint8 HWC BGR (cv2.imread layout), /255 folded into the stem kernel"),
                   "forward": "cal_hrnet_forward (C++ engine)" if os.environ.get("CAL_ENGINE", "1") != "0" else "hrnet.py schedule",
                   "camera_solve_inputs": ("synthetic keypoints of plausible cameras (keypoints_override): random-init weights "
                                           "give conf ~ 1/58 < every threshold; the networks' own decoded keypoints are still "
                                           "produced every step") if solve else None,
                   "l2": "working set larger than L2 (100 MB of uint8 frames + GBs of activations per step)",
                   "camera_solve_schedule": (("second stream, under the next batch's networks (co-resident blocks, "
                                              f"{pipe.solve_headroom} B shared-memory headroom)") if pipe.overlap_solve
                                             else "same stream, after the networks") if solve else None,
                   "parallelism": f"frame shards x{world}, one NCCL all-gather of the results" if world > 1 else "single GPU"},
        "e2e": {"value": frames_total / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "h2d_copy_ms_alone": h2d_ms, "h2d_gbps_alone": host.numel() * host.element_size() / h2d_ms / 1e6},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "decode": dec, "kernels_ms_per_step": kernels_ms,
        "multi_gpu_check": mg_check,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    import torch
    from tests import inputs as I
    cores = os.cpu_count() or 1
    path = CpuPath(args.workload, cores)
    n = args.cpu_frames
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(1, 1, *args.net_hw)))
    solve = WORKLOADS[args.workload][2]
    synth = synthetic_keypoints(n + 1, seed=5) if solve else None
    path(x, None if synth is None else synth[:1])            # warm-up frame
    t0 = time.perf_counter()
    for i in range(n):
        path(x, None if synth is None else synth[i + 1:i + 2])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} frames of the same workload, batch 1, after 1 warm-up frame ({dt:.1f} s of CPU work)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CAL_BENCH_WORKLOAD", "full"), choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--height", type=int, default=H_IMG, help="network input height (BASELINE config 5 sweep: 540/720/1080)")
    ap.add_argument("--width", type=int, default=W_IMG)
    ap.add_argument("--cpu-frames", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--float-frames", action="store_true", help="feed float32 CHW frames (398 MB per batch) instead of uint8 HWC")
    ap.add_argument("--shapes-out", default="", help="write per-(kernel, shape) device times of one profiled step")
    args = ap.parse_args()
    # the network runs at --height x --width; keypoints stay in 960x540 coordinates (the camera
    # solve's hard-coded IMG_SIZE, prediction.py:24, 622-629), as the decode's `size` argument does
    NET_H, NET_W = args.height, args.width
    args.net_hw = (NET_H, NET_W)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout: libraries that write to file descriptor 1 behind
    # Python's back (NCCL's version banner, for one) are sent to stderr, and Python's own stdout
    # keeps the original descriptor.
    sys.stdout.flush()
    _out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_out, "w", buffering=1)
    main()
    sys.stdout.flush()
