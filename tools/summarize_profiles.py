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
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.environ.get("CAL_PROF_DIR", os.path.join(ROOT, "profiles"))   # CAL_PROF_DIR: summarise on the GPU box into gpurun_out/
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


def shapes():
    """The NVTX-captured single launches of tools/ncu_shapes.py: one row per shape with counter-backed traffic and
    tensor-pipe numbers, and profiles/traffic.json for bench.py's roofline.traffic."""
    p = os.path.join(OUT, f"{tag}_shapes.ncu-rep")
    if not os.path.exists(p):
        return
    import json
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    names = ["BasicBlock 48 @135x240 (conv-ReLU-conv-add-ReLU in one kernel)", "3x3 48->48 @135x240", "3x3 48->48 @135x240 +residual", "3x3 96->96 @68x120", "3x3 96->96 @68x120 +residual",
             "3x3 192->192 @34x60", "3x3 384->384 @17x30", "1x1 64->256 @135x240 +residual"]
    # algorithmic bytes per launch at batch 64: input + output (+ residual) in fp16 NHWC with the padded channel count
    B = 64
    alg = [B * 135 * 240 * 64 * 2 * 2, B * 135 * 240 * 64 * 2 * 2, B * 135 * 240 * 64 * 2 * 3, B * 68 * 120 * 128 * 2 * 2, B * 68 * 120 * 128 * 2 * 3,
           B * 34 * 60 * 192 * 2 * 2, B * 17 * 30 * 384 * 2 * 2, B * 135 * 240 * (64 + 256 + 256) * 2]
    flop = [4 * B * 135 * 240 * 48 * 48 * 9] + [2 * B * 135 * 240 * 48 * 48 * 9] * 2 + [2 * B * 68 * 120 * 96 * 96 * 9] * 2 + [2 * B * 34 * 60 * 192 * 192 * 9,
            2 * B * 17 * 30 * 384 * 384 * 9, 2 * B * 135 * 240 * 64 * 256]
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = lambda r, c: r[hdr.index(c)]
    out = []
    with open(os.path.join(PROF, f"{tag}_shapes_ncu_full.csv"), "w") as f:
        f.write("# ncu --set full --clock-control none --nvtx --nvtx-include cap/ : one warm launch of each dominant conv shape at batch 64 "
                "(tools/ncu_shapes.py); algorithmic bytes = input + output (+ residual) once, fp16 NHWC, padded channels (the fused BasicBlock: block input + block output; the two launches it replaces: 530.8 + 796.3 MB)\n")
        f.write("shape,kernel,duration_us,dram_read_MB,dram_write_MB,algorithmic_MB,traffic_over_algorithmic,dram_pct_of_peak,"
                "tensor_pipe_pct_elapsed,algorithmic_TFLOPs_per_s,registers,smem_dynamic_KB\n")
        for i, r in enumerate(rows[2:]):
            if i >= len(names):
                break
            scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}
            rd = float(col(r, "dram__bytes_read.sum")) * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(col(r, "dram__bytes_write.sum")) * scale[units[hdr.index("dram__bytes_write.sum")]]
            du = float(col(r, "gpu__time_duration.sum"))
            du_us = du * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[hdr.index("gpu__time_duration.sum")]]
            a_mb = alg[i] / 1e6
            f.write(f"{names[i]},{short(col(r, 'Kernel Name'))},{du_us:.1f},{rd:.1f},{wr:.1f},{a_mb:.1f},{(rd + wr) / a_mb:.3f},"
                    f"{col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')},"
                    f"{col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')},{flop[i] / du_us / 1e6:.1f},"
                    f"{col(r, 'launch__registers_per_thread')},{col(r, 'launch__shared_mem_per_block_dynamic')}\n")
            out.append({"shape": names[i], "kernel": short(col(r, "Kernel Name")), "duration_us_under_ncu": du_us,
                        "dram_bytes": (rd + wr) * 1e6, "algorithmic_bytes": alg[i]})
    json.dump({"source": f"profiles/{tag}_shapes_ncu_full.csv", "launches": out}, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, f"{tag}_shapes_ncu_full.csv")).read())


launches()
shapes()
for n in sys.argv[3:] or ["halo", "convtc", "head", "combine", "decode", "linedecode", "solve", "stem", "conv"]:
    full(n)
import glob
extra = [os.path.basename(q) for q in glob.glob(os.path.join(OUT, f"{tag}_bench_*.json")) + glob.glob(os.path.join(OUT, f"{tag}_shapes_*.csv")) + glob.glob(os.path.join(OUT, f"{tag}_smoke.log"))]
for fn in sorted(set([f"{tag}_bench_{wl}.json", f"{tag}_bench_{wl}_ref.json", f"{tag}_smi.txt", f"{tag}_pytest_gpu.log"] + extra)):
    if os.path.exists(os.path.join(OUT, fn)):
        shutil.copy(os.path.join(OUT, fn), os.path.join(PROF, fn))
