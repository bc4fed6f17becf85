"""One pass of the full pipeline (both nets, decodes, line points, camera solve) for ncu captures:
python tools/ncu_pipeline.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccernet_calibration_sportlight_b200.pipeline import CalibrationPipeline  # noqa: E402
from tests import camera_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pipe = CalibrationPipeline("cuda:0", workload="full")
x = torch.rand(B, 3, 540, 960, device="cuda:0")
kp = torch.from_numpy(camera_inputs.synthetic_predictions(B, seed=100)).cuda()
for _ in range(2):
    out = pipe(x, keypoints_override=kp)
torch.cuda.synchronize()
print("done", {k: tuple(v.shape) for k, v in out.items()})
