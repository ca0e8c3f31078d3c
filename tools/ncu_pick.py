"""Prints selected metrics from an `ncu --page raw --csv` export (one column per launch)."""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
extra = sys.argv[2:]
for i, h in enumerate(hdr):
    if h in WANT or any(e in h for e in extra):
        print("%-90s %-10s %s" % (h, units[i], [r[i] for r in data]))
