"""Where does the end-to-end loop lose time against the device-resident loop? (profiling experiment)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from soccernet_calibration_sportlight_b200.pipeline import CalibrationPipeline
from tests import camera_inputs
B, K = 64, 10
pipe = CalibrationPipeline("cuda:0", workload="full")
g = torch.Generator().manual_seed(0)
host = torch.randint(0, 256, (B, 3, 540, 960), generator=g, dtype=torch.uint8).float().div_(255.0).pin_memory()
frames = host.cuda()
kp = torch.from_numpy(camera_inputs.synthetic_predictions(B, seed=100)).cuda()
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name:60s} {dt / K * 1e3:7.1f} ms/step  {B * K / dt:6.0f} frames/s", flush=True)
def resident():
    for _ in range(K): pipe(frames, keypoints_override=kp)
def resident_d2h():
    for _ in range(K): pipe(frames, keypoints_override=kp)["cameras"].to("cpu")
def stream(to_host):
    for r in pipe.run_stream((host for _ in range(K)), keypoints_override=kp, to_host=to_host): pass
for _ in range(3): resident()
for rep in range(2):
    timed("device-resident frames, no read-back", resident)
    timed("device-resident frames, result read back every step", resident_d2h)
    timed("run_stream: pinned host frames, results stay on device", lambda: stream(False))
    timed("run_stream: pinned host frames, result read back every step", lambda: stream(True))
