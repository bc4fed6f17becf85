"""Turns the scratch ncu outputs under gpurun_out/ into the small tracked summaries under
profiles/:  python tools/summarize_profiles.py <tag> [workload]"""
import csv
import collections
import io
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
wl = sys.argv[2] if len(sys.argv) > 2 else "kp_decode"
os.makedirs(PROF, exist_ok=True)

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg"]


def short(name):
    m = re.search(r"(\w+)(<[^>]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def launches():
    p = os.path.join(OUT, f"{tag}_launches_{wl}.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(l for l in open(p) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        d = agg.setdefault(short(r[ik]), [0, 0.0])
        d[0] += 1
        d[1] += float(r[iv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches_{wl}_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --workload %s --steps 2 --warmup 3`\n" % wl)
        f.write("# cold-cache, serialised per-launch times: compare SHARES with bench.py's kernels_ms_per_step, not absolutes\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{ns / 1e6:.3f},{ns / tot:.4f}\n")
    print(open(os.path.join(PROF, f"{tag}_launches_{wl}_summary.csv")).read())


def full(name):
    p = os.path.join(OUT, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(p):
        return
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    cols = [c for c in KEEP if c in hdr]
    with open(os.path.join(PROF, f"{tag}_{name}_ncu_full.csv"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; selected columns of {tag}_{name}.ncu-rep\n")
        w = csv.writer(f)
        w.writerow([c + (f" [{units[hdr.index(c)]}]" if units[hdr.index(c)] else "") for c in cols])
        for r in rows[2:]:
            w.writerow([short(r[hdr.index(c)]) if c == "Kernel Name" else r[hdr.index(c)] for c in cols])
    print(open(os.path.join(PROF, f"{tag}_{name}_ncu_full.csv")).read())


launches()
for n in sys.argv[3:] or ["halo", "convtc", "head", "combine", "decode", "linedecode", "solve", "stem", "conv"]:
    full(n)
for fn in (f"{tag}_bench_{wl}.json", f"{tag}_bench_{wl}_ref.json", f"{tag}_smi.txt", f"{tag}_pytest_gpu.log"):
    if os.path.exists(os.path.join(OUT, fn)):
        shutil.copy(os.path.join(OUT, fn), os.path.join(PROF, fn))
