"""Per-kernel roofline table of one ATV step from the committed `ncu --set full` raw pages (profiles/r02/r02d_*_raw.csv):
time, DRAM bytes, achieved DRAM rate as a fraction of the measured copy bandwidth (MEASURED_PEAKS.json), tensor-pipe activity.
    python profiles/roofline_table.py > profiles/r02/roofline_table_d.md"""
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = PK["hbm_gbs"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def rows(name):
    p = os.path.join(ROOT, "profiles", "r02", f"r02d_{name}_raw.csv")
    r = list(csv.reader(open(p)))
    return r[0], r[1], r[2:]


def val(h, u, d, k):
    i = h.index(k)
    return float(d[i].replace(",", "")) * UNIT.get(u[i], 1)


print(f"# Roofline view of one HybridRNN_ATV step (final round-2 state)\n\nFrom the `ncu --set full --clock-control none` raw pages "
      f"`profiles/r02/r02d_{{mab,attn,gemm,gemmw}}_raw.csv` (per launch; cold cache, serialised).  HBM peak = measured copy "
      f"bandwidth {HBM:.0f} GB/s (`MEASURED_PEAKS.json`); `DRAM frac` = (dram read + write bytes) / time / peak; `tensor %` = "
      f"`sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active`.  The row-wise kernels (`dln_*`, `colsum_*`) are in "
      f"`profiles/r01n_ncu_summary.md` (their capture of this round was cut off).\n")
tot = {}
for grp in ("mab", "attn", "gemm", "gemmw"):
    h, u, data = rows(grp)
    print(f"## {grp}\n\n| kernel | grid | ms | DRAM MB | GB/s | DRAM frac | tensor % | issue % |\n|---|---|---|---|---|---|---|---|")
    for d in data:
        name = re.sub(r"\(.*", "", d[h.index("Kernel Name")]).replace("void ", "").replace("lsthm::", "")
        ms = val(h, u, d, "gpu__time_duration.sum")
        by = val(h, u, d, "dram__bytes_read.sum") + val(h, u, d, "dram__bytes_write.sum")
        gbs = by / (ms * 1e-3) / 1e9
        tp = val(h, u, d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        ia = val(h, u, d, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        print(f"| {name} | {d[h.index('launch__grid_size')]} | {ms:.3f} | {by / 1e6:.0f} | {gbs:.0f} | {gbs / HBM:.2f} | {tp:.1f} | {ia:.1f} |")
        a = tot.setdefault(grp, [0.0, 0.0, 0.0])
        a[0] += ms; a[1] += by; a[2] += tp * ms
    print()
print("## per family\n\n| family | launches' time [ms] | DRAM GB | mean GB/s | DRAM frac | time-weighted tensor % |\n|---|---|---|---|---|---|")
for grp, (ms, by, tpw) in tot.items():
    gbs = by / (ms * 1e-3) / 1e9
    print(f"| {grp} | {ms:.2f} | {by / 1e9:.2f} | {gbs:.0f} | {gbs / HBM:.2f} | {tpw / ms:.1f} |")
