"""Builds profiles/<tag>_ncu_summary.md (+ dram_traffic.json) from the CSV pages exported ON THE GPU BOX
(`ncu -i X.ncu-rep --page raw --csv`, `--page source --csv -k regex:<kernel>`): the .ncu-rep files of a full
step exceed gpurun's 64 MiB return limit, so only the CSV pages travel back.
    python profiles/summarize_ncu_csv.py r01f
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0,
        "nsecond": 1e-6, "second": 1e3}


def raw_rows(name):
    p = os.path.join(G, f"{tag}_{name}_raw.csv")
    if not os.path.exists(p):
        return [], [], []
    rows = list(csv.reader(open(p)))
    return rows[0], rows[1], rows[2:]


def val(hdr, units, d, key):
    if key not in hdr:
        return None
    i = hdr.index(key)
    try:
        return float(d[i].replace(",", "")) * UNIT.get(units[i], 1)
    except ValueError:
        return None


def segments(kernel):
    p = os.path.join(G, f"{tag}_{kernel}_src.csv")
    if not os.path.exists(p):
        return []
    srows = list(csv.reader(open(p)))
    h = next((r for r in srows if "Source" in r and "Address" in r), None)
    if not h:
        return []
    si, sa = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    body = [r for r in srows[srows.index(h) + 1:] if len(r) > sa and r[sa].isdigit()]
    tot = sum(int(r[sa]) for r in body) or 1
    out, acc, fm, tc = [], 0, 0, 0
    for r in body:
        acc += int(r[sa]); fm += ("FFMA" in r[si]); tc += ("UTC" in r[si] and "MMA" in r[si])
        if "BAR.SYNC" in r[si] or "EXIT" in r[si]:
            out.append((100 * acc / tot, fm, tc)); acc = fm = tc = 0
    return out


lines = [f"# ncu summary `{tag}`", "",
         "Source: `ncu --set full --clock-control none [--import-source on] -k regex:<kernel>` on `bench.py --steps 1..2 --warmup 1..3 "
         "--no-e2e --no-cpu-baseline` (`--model sps --steps 1` for the sps cell; from r01n on: profiles/capture_ncu.sh), 1 GPU (B200); each run was preceded by the same "
         "command without ncu exiting 0.  The raw/source pages were exported to CSV on the GPU box "
         "(the .ncu-rep files exceed the 64 MiB return limit).  Numbers are per launch; durations under ncu are cold-cache "
         "and serialised.", ""]
traffic = {}
tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
if os.path.exists(tpath):
    traffic = json.load(open(tpath))
for grp in ("mab", "sps", "attn"):
    hdr, units, data = raw_rows(grp)
    for d in data:
        name = re.sub(r"\(.*", "", d[hdr.index("Kernel Name")]).replace("void ", "")
        short = name.split("<")[0].split("::")[-1]
        lines += [f"## {name}", "", "| metric | value |", "|---|---|"]
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} [{units[i]}] | {d[i]} |")
        rd, wr = val(hdr, units, d, "dram__bytes_read.sum"), val(hdr, units, d, "dram__bytes_write.sum")
        if rd is not None:
            traffic[short] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                              "duration_ms_under_ncu": val(hdr, units, d, "gpu__time_duration.sum"),
                              "source": f"profiles/{tag}_ncu_summary.md"}
            lines += ["", f"DRAM traffic per launch: read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB"]
        seg = segments(short)
        if seg:
            lines += ["", "Warp-sample share per barrier-delimited segment (SASS order = stage order; barrier-wait samples land in the "
                      "segment after the BAR):", "", "| seg | share | FFMA/FFMA2 instrs | UMMA instrs |", "|---|---|---|---|"]
            lines += [f"| {i} | {s:.1f}% | {f} | {t} |" for i, (s, f, t) in enumerate(seg)]
        lines.append("")
for grp, title in (("gemm", "lsthm::gemm3_kernel (general operands; the weight-gradient TN products)"),
                   ("gemmw", "lsthm::gemm3w_kernel (weight-stationary NT / NN products)"),
                   ("dln", "row-wise HBM-bound kernels (fused dropout+residual+LayerNorm, column sums)")):
  hdr, units, data = raw_rows(grp)
  if data:
    lines += [f"## {title} — all launches of one step", "",
              "| variant | grid | ms | tensor pipe active % | issue active % | DRAM thr % | L2 thr % | DRAM MB (r+w) |", "|---|---|---|---|---|---|---|---|"]
    tot = 0.0
    for d in data:
        name = re.sub(r"\(.*", "", d[hdr.index("Kernel Name")]).replace("void ", "").replace("lsthm::", "")
        g = lambda k: val(hdr, units, d, k)
        ms = g("gpu__time_duration.sum"); tot += ms
        lines.append(f"| {name} | {d[hdr.index('launch__grid_size')]} | {ms:.3f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {(g('dram__bytes_read.sum') + g('dram__bytes_write.sum')) / 1e6:.0f} |")
    lines += ["", f"Sum of the {len(data)} GEMM launches of one step: {tot:.2f} ms (under ncu).", ""]
json.dump(traffic, open(tpath, "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
print("wrote", f"profiles/{tag}_ncu_summary.md", list(traffic))
