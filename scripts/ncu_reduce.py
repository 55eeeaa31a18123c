"""Reduce `ncu --page raw --csv` output to the columns the roofline needs (one row per launch)."""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
# the first row with "ID" is the header, the next one the units
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
header, units, data = rows[h], rows[h + 1], rows[h + 2:]
tensor_cols = [c for c in header if "tensor" in c and c not in KEEP]
cols = [c for c in KEEP + tensor_cols if c in header]
idx = [header.index(c) for c in cols]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[i] for i in idx])
    for r in data:
        if len(r) >= len(header):
            w.writerow([r[i] for i in idx])
print("reduced %d launches, %d columns" % (len(data), len(cols)))
