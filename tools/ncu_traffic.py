"""DRAM traffic of the conv kernels per step from an ncu metrics CSV (dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum over `tools/profile_layers.py`, which runs `passes` identical eager steps):
python tools/ncu_traffic.py gpurun_out/conv_traffic.csv 3 > profiles/conv_dram_traffic.json"""
import csv, json, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6}
launches = {}
for r in rows:
    if not any(t in r[ki] for t in ("conv_", "dwpw_kernel", "stem_pair_kernel", "bneck_pair_kernel")):
        continue
    d = launches.setdefault(int(r[ii]), {"name": re.sub(r"\(.*", "", r[ki]).replace("specyolo::", "").replace("void ", "")})
    d[r[mi]] = float(r[vi].replace(",", "")) * unit.get(r[ui], 1.0)
ids = sorted(launches)
per = len(ids) // passes
last = ids[-per:]
rd = sum(launches[i].get("dram__bytes_read.sum", 0.0) for i in last)
wr = sum(launches[i].get("dram__bytes_write.sum", 0.0) for i in last)
t = sum(launches[i].get("gpu__time_duration.sum", 0.0) for i in last)
print(json.dumps({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum on tools/profile_layers.py 64 (last of %d eager passes)" % passes,
                  "conv_launches_per_step": per, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
                  "dram_bytes_per_step": rd + wr, "conv_time_ns_under_ncu": t}, indent=1))
