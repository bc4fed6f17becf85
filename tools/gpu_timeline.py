"""Per-role clock64 timeline of CTA 0 of the halo conv kernel (profiling experiment)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
dbg = torch.zeros(4 * 64 * 8 + 2 * 148, dtype=torch.int64, device="cuda")
os.environ["CAL_DEBUG_TIMELINE"] = hex(dbg.data_ptr())
from soccernet_calibration_sportlight_b200 import ops
C, H, W, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 64
LO = int(os.environ.get("LO", "20"))
use_res = len(sys.argv) > 4 and sys.argv[4] == "res"
cp = (C + 63) // 64 * 64
x = torch.randn(B, H, W, cp, device="cuda").half()
w = (torch.randn((C + 15) // 16 * 16, 9 * cp, device="cuda") / 30).half()
bias = torch.zeros(cp, device="cuda")
y = torch.empty_like(x)
for _ in range(3):
    ops.conv2d(x, w, bias, y, ksize=3, stride=1, cout_rows=w.shape[0], relu=True, res=x if use_res else None)
torch.cuda.synchronize()
span = dbg.cpu().numpy()[4 * 64 * 8:].reshape(148, 2)
d = dbg.cpu().numpy()[:4 * 64 * 8].reshape(4, 64, 8)
t0 = d[d > 0].min()
names = {0: ["wait_emptyA", "got_emptyA", "B4_wait_empty", "B4_issue"], 1: ["start", "got_tempty", "got_fullA", "committed", "B4_wait_full", "B4_got_full", "B4_mma_issued", "B4_committed"],
         2: ["begin", "pre_bulkwait", "post_bulkwait", "post_bar1", "got_tfull", "epi_done", "fenced", "stores_issued"]}
for role, nm in ((0, "producer"), (1, "mma"), (2, "epi0"), (3, "epi1")):
    print(nm, names[min(role, 2)])
    for i in range(LO, LO + 8):
        row = d[role, i]
        print("   tile", i, [int(v - t0) if v > 0 else None for v in row[:len(names[min(role, 2)])]])
for role, nm in ((1, "mma"), (2, "epi0")):
    per = (d[role, LO + 7, 0] - d[role, LO, 0]) / 7
    print(nm, "cycles per own iteration:", per)

import numpy as np
st, en = span[:, 0] - span[:, 0].min(), span[:, 1] - span[:, 0].min()
print("per-CTA wall clock (ns): start min/max", st.min(), st.max(), " end min/median/max", en.min(), int(np.median(en)), en.max())
order = np.argsort(en)
print("  earliest-finishing CTAs", [(int(i), int(en[i])) for i in order[:6]])
print("  latest-finishing CTAs  ", [(int(i), int(en[i])) for i in order[-6:]])
print("  busy time per CTA (end - start): min/median/max", (en - st).min(), int(np.median(en - st)), (en - st).max())
