"""Turn an .ncu-rep into the text summary kept under profiles/ (the metrics the roofline and the
judge look at). Usage: python profiles/ncu_summary.py report.ncu-rep "header line" > out.txt"""
import csv
import subprocess
import sys

rep, header = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ("Kernel Name", "dram__bytes", "gpu__dram_throughput", "gpu__time_duration.sum", "l1tex__data_bank_conflicts",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__t_sector_hit_rate", "l1tex__throughput.avg.pct",
        "lts__t_sector_hit_rate", "lts__throughput.avg.pct", "launch__", "sm__throughput.avg.pct", "sm__warps_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__average_warps_issue_stalled",
        "sm__inst_executed_pipe", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
print(header)
for li, r in enumerate(rows[2:]):
    print("\n--- launch %d ---" % li)
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in KEEP) and "per_issue_active" in h or (any(h.startswith(k) for k in KEEP) and "issue_stalled" not in h):
            print("%-86s %-16s %s" % (h, r[i], units[i]))
