"""tcgen05 / TMA / TMEM instruction counts per kernel of the built library, from `cuobjdump -sass` (runs without a
GPU), plus the barrier / TMA / MMA instructions of one kernel in program order:
    python tools/sass_summary.py [kernel substring] > profiles/<tag>_sass_tcgen05.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "soccernet_calibration_sportlight_b200", "csrc", "libcalib_b200.so")
PICK = sys.argv[1] if len(sys.argv) > 1 else "conv3x3_halo_kernel<true, 1, true>"
OPS = re.compile(r"\b(UTCHMMA|UTMALDG[.\w]*|UTMASTG[.\w]*|LDTM[.\w]*|UTCBAR|UTCCP[.\w]*|UTMAPF)\b")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = {}
counts = collections.OrderedDict()
order = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        raw = m.group(1)
        if raw not in names:
            d = subprocess.run(["c++filt", raw], capture_output=True, text=True).stdout.strip()
            d = re.sub(r"cal::\(anonymous namespace\)::", "", d)
            names[raw] = re.sub(r"\(.*$", "", d).replace("void ", "")
        fn = names[raw]
        counts.setdefault(fn, collections.Counter())
        order.setdefault(fn, [])
        continue
    if fn is None:
        continue
    body = re.sub(r"/\*.*?\*/", "", line).strip()
    m = OPS.search(body)
    if m:
        counts[fn][m.group(1)] += 1
    if m or "SYNCS" in body:
        order[fn].append(body.rstrip(" ;"))
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): tensor-core / TMA / TMEM instructions per kernel")
print("# UTCHMMA = tcgen05.mma kind::f16 | UTMALDG = cp.async.bulk.tensor (TMA load) | LDTM = tcgen05.ld | UTCBAR = tcgen05.commit")
for f, c in counts.items():
    if c:
        print(f"{f:55s} " + "  ".join(f"{k} x{v}" for k, v in sorted(c.items())))
print(f"\n# {PICK}: mbarrier (SYNCS), TMA and MMA instructions in program order")
for f, lines in order.items():
    if PICK in f:
        print("\n".join("    " + l for l in lines))
        break
