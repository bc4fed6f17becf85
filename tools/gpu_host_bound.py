"""Is the pipeline host-launch-bound?  Host enqueue time vs device time of one step, and the same
step replayed from a CUDA graph (profiling experiment)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from soccernet_calibration_sportlight_b200.pipeline import CalibrationPipeline
B = int(os.environ.get("B", "64"))
wl = os.environ.get("WL", "full")
pipe = CalibrationPipeline("cuda:0", workload=wl)
g = torch.Generator().manual_seed(0)
frames = torch.rand(B, 3, 540, 960, generator=g).cuda()
from tests import camera_inputs
kp_over = torch.from_numpy(camera_inputs.synthetic_predictions(B, seed=100)).cuda() if wl != "kp_decode" else None
def step():
    return pipe(frames, keypoints_override=kp_over)
for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"eager: host enqueue {1e3 * (t1 - t0):.1f} ms, until device done {1e3 * (t2 - t0):.1f} ms", flush=True)
try:
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        out = step()
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"graph replay: {e0.elapsed_time(e1) / 5:.1f} ms per step ({B / (e0.elapsed_time(e1) / 5) * 1e3:.0f} frames/s)", flush=True)
    ref = step(); torch.cuda.synchronize()
    graph.replay(); torch.cuda.synchronize()
    for k in ref:
        same = torch.equal(torch.nan_to_num(ref[k].double()), torch.nan_to_num(out[k].double()))
        print(f"  {k}: graph == eager: {same}")
except Exception as e:  # noqa
    import traceback; traceback.print_exc()
