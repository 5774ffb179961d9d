"""Summarise an .ncu-rep: per-kernel key metrics (raw page) — python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keys = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
        ("smsp__inst_executed.sum", "inst")]
for r in data:
    name = r[idx["Kernel Name"]][:60]
    parts = []
    for k, short in keys:
        if k in idx:
            v = r[idx[k]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{units[idx[k]] if short in ('time','dram_rd','dram_wr') else ''}")
    print(name, " ".join(parts))
