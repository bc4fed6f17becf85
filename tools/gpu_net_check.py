"""Whole-network parity + first timing on the GPU box.
    python tools/gpu_net_check.py [--full] [--bench B]
Compares the CUDA HRNet (fp16 storage, fp32 accumulate) with the fp32 CPU oracle on the
same seeded weights and frames; logs to gpurun_out/net_check.log."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_ref as O  # noqa: E402
from soccernet_calibration_sportlight_b200 import hrnet as P  # noqa: E402
from tests import inputs as I  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "net_check.log"), "a")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def compare(kind, H, W, B=1, seed=11):
    oracle = O.make_model(kind, seed=seed)
    net = P.HRNetHeatmap(P.w48_config(kind)).load_state_dict(oracle.state_dict()).to("cuda:0")
    x = torch.from_numpy(I.frames_to_tensor(I.frames_u8(21, B, H, W)))
    t0 = time.time()
    with torch.no_grad():
        ref = oracle(x)[-1]
    t_cpu = time.time() - t0
    got = net(x.cuda())[-1]
    torch.cuda.synchronize()
    got = got.cpu()
    d = (got - ref).abs()
    if kind == "keypoints":
        pr, pg = ref.exp(), got.exp()
    else:
        pr, pg = ref, got
    amax_r = pr.flatten(2).argmax(-1)
    amax_g = pg.flatten(2).argmax(-1)
    log(f"{kind} {H}x{W} B{B}: out {tuple(got.shape)} finite={bool(torch.isfinite(got).all())} "
        f"max_abs={float(d.max()):.4e} mean_abs={float(d.mean()):.4e} ref_range=[{float(ref.min()):.3f},{float(ref.max()):.3f}] "
        f"prob_max_abs={float((pr - pg).abs().max()):.4e} rel_l2={float((got - ref).norm() / ref.norm()):.4e} "
        f"argmax_agree={float((amax_r == amax_g).float().mean()):.3f} cpu_oracle_s={t_cpu:.2f}")
    return float(d.max())


def bench(kind, B, H=540, W=960, iters=3):
    net = P.HRNetHeatmap(P.w48_config(kind)).to("cuda:0")
    x = torch.rand(B, 3, H, W, device="cuda:0")
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        net(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gf = 507.82 if kind == "keypoints" else 371.4
    log(f"bench {kind} B{B} {H}x{W}: {ms:.2f} ms/forward  {B / ms * 1e3:.1f} frames/s  "
        f"{gf * B / ms:.1f} algorithmic TFLOP/s  mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--bench", type=int, default=0)
    a = ap.parse_args()
    log(f"=== net_check {time.strftime('%F %T')} ===")
    compare("keypoints", 96, 160)
    compare("lines", 96, 160)
    if a.full:
        compare("keypoints", 540, 960)
        compare("lines", 540, 960)
    if a.bench:
        for kind in ("keypoints", "lines"):
            bench(kind, 8)
            bench(kind, a.bench)
