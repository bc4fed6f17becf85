"""Golden vectors for the two heat-map decoders, produced by the UNMODIFIED
reference (build container only):
    python tests/golden/make_golden_decode.py
Reference entry points run here:
  src/models/hrnet/transforms.py:224-239  HRNetPredictionTransform
  src/models/line/transforms.py:224-280   EHMPredictionTransform.mask_heat_points_gauss
  src/models/hrnet/loss.py:21-87          create_target (the authors' 'perfect heat-map'
                                          debug hook, metamodel.py:69-75)
Writes tests/golden/decode_keypoints.npz and decode_lines.npz and prints how the
oracle restatement (oracle/decode_ref.py) agrees with the reference."""
import os
import sys

import numpy as np
import torch

import refimport

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
refimport.setup()
from src.models.hrnet.loss import HRNetLoss  # noqa: E402
from src.models.hrnet.transforms import HRNetPredictionTransform  # noqa: E402
from src.models.line.transforms import EHMPredictionTransform  # noqa: E402

from oracle import decode_ref as O  # noqa: E402
from tests import inputs as I  # noqa: E402

torch.set_num_threads(8)
out = {}
report = []


def run_kp(name, logp, size, store_input=False):
    ref = HRNetPredictionTransform(size)(torch.from_numpy(logp)).numpy()
    mine = O.keypoint_decode_np(logp, size)
    lit = O.keypoint_decode_torch(torch.from_numpy(logp), size).numpy()
    assert np.array_equal(ref, lit), name
    idx_ok = np.array_equal(ref[..., :2], mine[..., :2])
    conf_ulp = np.abs(ref[..., 2].view(np.int32) - mine[..., 2].view(np.int32)).max()
    report.append(f"kp  {name:28s} shape={logp.shape} idx_equal={idx_ok} conf_max_ulp={conf_ulp}")
    out[f"{name}__out"] = ref
    out[f"{name}__size"] = np.array(size)
    if store_input:
        out[f"{name}__in"] = logp
    return idx_ok


# (1) generated-by-formula inputs (rebuilt bit-identically by tests/inputs.py)
run_kp("hashed_small", I.hashed_logp(1, (2, 58, 20, 24)), (40, 48))
run_kp("hashed_ragged", I.hashed_logp(2, (1, 5, 7, 13)), (21, 39))
run_kp("hashed_full", I.hashed_logp(3, (1, 58, 270, 480)), (540, 960))
run_kp("gauss_small", I.gaussian_logp(4, 2, 58, 20, 24), (40, 48))
run_kp("gauss_full", I.gaussian_logp(5, 1, 58, 270, 480), (540, 960))
run_kp("gauss_720p", I.gaussian_logp(6, 1, 58, 360, 640), (540, 960))
# (2) the reference's own target construction -> log (inputs stored: transcendental)
loss = HRNetLoss(sigma=3.0, stride=2, pred_size=(34, 60))
g = torch.Generator().manual_seed(0)
kp = torch.rand(2, 57, 2, generator=g) * torch.tensor([59.0, 33.0])
kp = torch.cat([kp, torch.ones(2, 57, 1)], -1)
kp[0, 5] = 0.0   # invisible
kp[1, 11] = 0.0
kp[0, 7, :2] = torch.tensor([10.5, 20.5])      # half-pixel centre: 4-way tie
target = loss.create_target(kp[..., :3].clone())
with np.errstate(divide="ignore"):
    logt = torch.log(target).numpy()
run_kp("target_log", logt, (68, 120), store_input=True)
# (3) log_softmax of random normal logits (stored)
z = torch.randn(2, 58, 16, 20, generator=g) * 3
run_kp("logsoftmax_rand", torch.log_softmax(z, 1).numpy(), (32, 40), store_input=True)
np.savez_compressed(os.path.join(HERE, "decode_keypoints.npz"), **out)

out = {}


def run_line(name, heat, sigma, store_input=False):
    ref = EHMPredictionTransform.mask_heat_points_gauss(torch.from_numpy(heat), sigma=sigma).numpy()
    mine = O.line_decode_np(heat, sigma)
    idx_ok = np.array_equal(ref[..., :2], mine[..., :2])
    rel = np.abs(ref[..., 2] - mine[..., 2]).max()
    report.append(f"line {name:27s} shape={heat.shape} sigma={sigma} idx_equal={idx_ok} val_max_abs={rel:.3e}")
    out[f"{name}__out"] = ref
    out[f"{name}__sigma"] = np.array(float(sigma))
    if store_input:
        out[f"{name}__in"] = heat
    # scaled transform (line/transforms.py:216-222)
    t = EHMPredictionTransform(scale=4, sigma=sigma)(torch.from_numpy(heat)).numpy()
    out[f"{name}__scaled4"] = t


run_line("tent_small_s3", I.two_peak_heat(7, 2, 23, 17, 30), 3.0)
run_line("tent_small_s6", I.two_peak_heat(8, 1, 23, 17, 30), 6)
run_line("tent_full_s3", I.two_peak_heat(9, 1, 23, 135, 240), 3.0)
run_line("tent_ragged", I.two_peak_heat(10, 1, 3, 5, 11), 2.5)
sm = torch.softmax(torch.randn(1, 23, 34, 60, generator=g) * 4, 1).numpy()
run_line("softmax_rand", sm, 3.0, store_input=True)
np.savez_compressed(os.path.join(HERE, "decode_lines.npz"), **out)
print("\n".join(report))
with open(os.path.join(HERE, "decode_report.txt"), "w") as f:
    f.write("\n".join(report) + "\n")
