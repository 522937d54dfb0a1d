"""Aggregates an ncu launch list (gpu__time_duration.sum per launch, CSV) into a per-kernel table for ONE bench step.
    python profiles/launch_breakdown.py gpurun_out/launches_<tag>.csv [anchor_kernel_substring]
A step = the launches between the last two occurrences of the anchor kernel (default: mab_fwd / sps_fwd)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
anchor = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "_fwd_kernel"
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr, data = rows[0], rows[1:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
idx = [i for i, d in enumerate(data) if anchor in d[ki] and ("mab" in d[ki] or "sps" in d[ki])]
per = 1
if "sps" in data[idx[-1]][ki]:
    per = 2   # two directions per step
seg = data[idx[-1 - per]:idx[-1]]
agg = collections.OrderedDict()
for d in seg:
    n = re.sub(r"\(.*", "", d[ki]).replace("void ", "")
    n = re.sub(r"at::native::|<unnamed>::|at::", "", n)[:100]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(d[vi]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"one step: {len(seg)} launches, {tot:.3f} ms summed kernel time (ncu: serialised, cold cache)\n")
print("| ms | % | launches | kernel |\n|---:|---:|---:|---|")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {v[1]:.3f} | {100 * v[1] / tot:.1f} | {v[0]} | `{n}` |")
if "--detail" in sys.argv:
    print()
    for d in seg:
        if "gemm3" in d[ki] and "reduce" not in d[ki] and "pack" not in d[ki] or "attn" in d[ki]:
            print(f"{float(d[vi]) / 1e3:9.1f} us  grid {d[gi]:>16} {re.sub(r'.*lsthm::', '', d[ki])[:40]}")
