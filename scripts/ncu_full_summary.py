"""Summarise the launches of an `ncu --set full` report (no GPU needed): duration, DRAM bytes, pipe activity, launch shape.
Usage: python scripts/ncu_full_summary.py <report.ncu-rep> [title]"""
import csv
import io
import subprocess
import sys

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__cluster_dim_x", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
if len(sys.argv) > 2:
    print("## " + sys.argv[2] + "\n")
for r in rows[2:]:
    print("  Kernel Name = " + r[ix["Kernel Name"]])
    for k in KEYS:
        if k in ix:
            print(f"  {k} = {r[ix[k]]} {units[ix[k]]}")
    print()
