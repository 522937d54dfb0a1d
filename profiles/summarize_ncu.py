"""Turns the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.
    python profiles/summarize_ncu.py <tag>     # e.g. r01b  -> reads gpurun_out/prof_<tag>.ncu-rep, launches_<tag>.csv
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
lcsv = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
out_md = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(v) * mult


lines = ["# ncu summary `%s`" % tag, "",
         "Source: `ncu --set full --clock-control none --import-source on` on `bench.py --steps 2 --warmup 3 --no-e2e "
         "--no-cpu-baseline` (1 GPU, B200), report kept in gpurun_out/ (scratch); numbers below are per launch.", ""]
traffic = {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        name = re.sub(r"\(.*", "", d[hdr.index("Kernel Name")]).replace("void ", "")
        lines += [f"## {name}", "", "| metric | unit | value |", "|---|---|---|"]
        vals = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (d[i], units[i])
                lines.append(f"| {k} | {units[i]} | {d[i]} |")
        rd = to_bytes(*vals["dram__bytes_read.sum"]); wr = to_bytes(*vals["dram__bytes_write.sum"])
        key = "mab_fwd_kernel" if "fwd" in name else "mab_bwd_kernel" if "bwd" in name else name
        traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                        "duration_ms_under_ncu": float(vals["gpu__time_duration.sum"][0]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(vals["gpu__time_duration.sum"][1], 1),
                        "source": f"profiles/{tag}_ncu_summary.md"}
        lines += ["", f"DRAM traffic per launch: read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB", ""]
        # per-barrier-segment stall sampling (SASS order), tells which stage of a step the time goes to
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + name.split("::")[-1].split("<")[0]],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        h = next((r for r in srows if "Source" in r and "Address" in r), None)
        if h:
            si, sa = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
            body = [r for r in srows[srows.index(h) + 1:] if len(r) > sa and r[sa].isdigit()]
            tot = sum(int(r[sa]) for r in body) or 1
            seg, acc, ff = 0, 0, 0
            lines += ["Warp-sample share per barrier-delimited segment (SASS order = stage order of one step; barrier-wait samples "
                      "land on the instruction after the BAR, i.e. in the *next* segment):", "", "| seg | share | FFMA2/FFMA | ",
                      "|---|---|---|"]
            for r in body:
                acc += int(r[sa]); ff += ("FFMA" in r[si])
                if "BAR.SYNC" in r[si] or "EXIT" in r[si]:
                    lines.append(f"| {seg} | {100 * acc / tot:.1f}% | {ff} |")
                    seg += 1; acc = 0; ff = 0
            lines.append("")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)

if os.path.exists(lcsv):
    ll = [l for l in open(lcsv) if not l.startswith("==")]
    rows = list(csv.DictReader(ll))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in rows:
        n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")
        n = re.sub(r"<.*", "", n)[:90]
        v = float(x["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(x["Metric Unit"], 1e-6)
        agg[n][0] += 1; agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    lines += [f"## launch list of one step ({len(rows)} launches, {tot:.2f} ms summed, ncu-serialised: compare shares)", "",
              "| ms | share | launches | kernel |", "|---|---|---|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        lines.append(f"| {t:.3f} | {100 * t / tot:.1f}% | {c} | `{n}` |")
    ours = sum(t for n, (c, t) in agg.items() if n.startswith("lsthm::"))
    lines += ["", f"Our kernels (`lsthm::*`): {ours:.3f} ms = {100 * ours / tot:.1f}% of the step; the rest is PyTorch/cuBLAS "
              "(encoders, hoisted input/weight-gradient GEMMs, head, loss).", ""]
open(out_md, "w").write("\n".join(lines) + "\n")
print(out_md, traffic)
