"""Barrier-delimited phase profile of a kernel from an `ncu --page source --csv` export: share of warp-stall samples,
share of executed instructions, top stall reasons and top opcodes per phase (phases are split at BAR instructions,
SASS order = stage order of the persistent step loop).   python profiles/phase_profile.py <src.csv> [min_share_pct]"""
import collections
import csv
import sys

rows_all = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
blocks, cur, hdr = [], None, None
for row in rows_all:
    if row and row[0] == "Kernel Name":
        cur = []
        blocks.append((row[1], cur))
    elif row and row[0] == "Address":
        hdr = row
    elif cur is not None and len(row) > 5:
        cur.append(row)
name, rows = blocks[0]
si, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot, ninst = sum(int(x[si]) for x in rows), sum(int(x[ii]) for x in rows)
print(f"{name[:70]}: {len(rows)} SASS lines, {tot} samples, {ninst} warp instructions")
stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
phase, agg = 0, collections.OrderedDict()
for x in rows:
    op = x[1].split()
    opn = op[1] if op[0].startswith("@") else op[0]
    a = agg.setdefault(phase, {"s": 0, "i": 0, "ops": collections.Counter(), "st": collections.Counter()})
    a["s"] += int(x[si]); a["i"] += int(x[ii]); a["ops"][opn.split(".")[0]] += int(x[ii])
    for c in stallcols:
        a["st"][hdr[c]] += int(x[c])
    if opn.startswith("BAR"):
        phase += 1
print("| phase | samples % | instr % | top stalls | top opcodes (M warp-instr) |\n|---|---|---|---|---|")
for p, a in agg.items():
    if a["s"] < tot * min_share / 100:
        continue
    top = ", ".join(f"{k[6:]} {v * 100 // max(1, a['s'])}%" for k, v in a["st"].most_common(3))
    ops = ", ".join(f"{k} {v / 1e6:.0f}" for k, v in a["ops"].most_common(5))
    print(f"| {p} | {a['s'] * 100 / tot:.1f} | {a['i'] * 100 / ninst:.1f} | {top} | {ops} |")
