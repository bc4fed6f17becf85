#!/bin/bash
TAG=r3g
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_abi.py tests/test_engine_gpu.py -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|^E   |FAILED|PASSED" $OUT/${TAG}_pytest_gpu.log | cut -c1-250 | tail -40
python - <<'PY'
import torch, numpy as np
from soccernet_calibration_sportlight_b200 import hrnet as P, ops
from tests import inputs as I
cfg = P.w48_config("keypoints")
a = P.HRNetHeatmap(cfg, kind="keypoints"); a.load_state_dict(P.init_state_dict(cfg, "keypoints", seed=3)); a.to("cuda:0")
fr = I.frames_u8(9, 2, 96, 160)
xf = torch.from_numpy(I.frames_to_tensor(fr)).contiguous().cuda(); x8 = torch.from_numpy(fr).cuda()
s1 = torch.empty((2,48,80,64), dtype=torch.float16, device="cuda"); s2 = torch.empty_like(s1)
ops.stem_conv(xf, a.stem_w, a.stem_b, s1); ops.stem_conv(x8, a.stem_w, a.stem_b, s2)
print("stem equal", torch.equal(s1, s2), float((s1.float()-s2.float()).abs().max()))
h1 = a(xf)[-1]; h2 = a(xf)[-1]; h3 = a(x8)[-1]
print("engine float twice equal", torch.equal(h1,h2), "float vs u8", torch.equal(h1,h3), float((h1-h3).abs().max()))
a.use_engine = False
p1 = a(xf)[-1]; p3 = a(x8)[-1]
print("python float vs engine float", torch.equal(p1,h1), float((p1-h1).abs().max()), "python u8 vs python float", torch.equal(p1,p3))
PY
