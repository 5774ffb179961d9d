"""Top stall lines of an .ncu-rep source page: python tools/ncu_source.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        h = i; break
hdr = rows[h]; data = rows[h + 1:]
ix = {k: i for i, k in enumerate(hdr)}
si = ix["Warp Stall Sampling (All Samples)"]
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[si] or 0) for r in data if len(r) > si)
print("total samples", tot)
top = sorted([r for r in data if len(r) > si], key=lambda r: -int(r[si] or 0))[:N]
for r in top:
    reasons = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:3]
    print(f"{int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  {r[ix['Source']][:70]:70s} {reasons}")
