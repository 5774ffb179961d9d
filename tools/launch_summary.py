"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, total time, share.
python tools/launch_summary.py gpurun_out/launches.csv [steps_in_capture] > profiles/rNN_launches.md"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki]).replace("specyolo::", "").replace("void ", "")
    name = re.sub(r"<.*", "<>", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
own = {k: v for k, v in agg.items() if not k.startswith(("native::", "at::", "fold_pack", "cuda::"))}
tot = sum(v[1] for v in own.values())
print(f"# ncu launch list summary: {sys.argv[1]} ({len(rows)} launches captured, cold-cache serialised times: compare shares)\n")
print("| kernel | launches | total ms | share of libspecyolo time |\n|---|---|---|---|")
for k, (n, t) in sorted(own.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t/1e6:.3f} | {100*t/tot:.1f} % |")
oth = {k: v for k, v in agg.items() if k not in own}
print("\nOther kernels in the capture (weight folding at load time, torch fill/copy plumbing):\n")
for k, (n, t) in sorted(oth.items(), key=lambda kv: -kv[1][1]):
    print(f"* `{k}`: {n} launches, {t/1e6:.3f} ms")
