"""Device time of the camera solve on the bench's synthetic keypoints."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from soccernet_calibration_sportlight_b200 import ops
from tests import camera_inputs, camera_parity as CP
for B in (64, 512):
    kp = torch.from_numpy(camera_inputs.synthetic_predictions(B, seed=100)).cuda()
    P = CP.make_params("iterative_voter", 0.5)
    for _ in range(2):
        ops.camera_solve(kp, P)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rec = ops.camera_solve(kp, P)
    e1.record()
    torch.cuda.synchronize()
    valid = int((rec.view(torch.int32).reshape(B, 32)[:, 30] == 1).sum())
    print(f"B={B}: {e0.elapsed_time(e1) / 5:.2f} ms per launch, {valid}/{B} cameras", flush=True)
