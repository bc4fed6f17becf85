"""Short keypoint (+line) forward + decode for ncu captures: python tools/ncu_forward.py [B] [kind]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccernet_calibration_sportlight_b200 import hrnet as P, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = sys.argv[2] if len(sys.argv) > 2 else "keypoints"
net = P.HRNetHeatmap(P.w48_config(kind)).to("cuda:0")
x = torch.rand(B, 3, 540, 960, device="cuda:0")
for _ in range(2):
    heat = net(x)[-1]
    out = ops.kp_decode(heat, (540, 960)) if kind == "keypoints" else ops.line_decode(heat, 3.0)
torch.cuda.synchronize()
print("done", tuple(out.shape))
