"""Condense `ncu -i X.ncu-rep --page raw --csv` into the short per-kernel summary kept under
profiles/ (one "metric = value unit" line per metric, kernels separated by '---').

    python tools/ncu_summary.py gpurun_out/X_raw.csv "header comment" > profiles/ncu_X_summary.txt
"""
import csv
import re
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic|cluster.*)|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r".*pipe_tensor.*(dmma|cycles_active_realtime).*|.*sm__pipe_fp64_cycles_active_realtime.*|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
    r"dram__bytes\.sum\.per_second|l1tex__t_bytes_pipe_lsu_mem_global_op_(ld|st)\.sum)$")

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr_i], rows[hdr_i + 1]
kcol = names.index("Kernel Name")
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
first = True
for r in rows[hdr_i + 2:]:
    if len(r) != len(names):
        continue
    if not first:
        print("---")
    first = False
    print("Kernel Name = " + r[kcol])
    for n, u, val in sorted(zip(names, units, r)):
        if KEEP.match(n) and val != "":
            print("%s = %s %s" % (n, val, u))
