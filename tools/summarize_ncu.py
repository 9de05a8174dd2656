#!/usr/bin/env python
"""Summarise an `ncu --set full` report into a small CSV (one row per kernel launch)
with the metrics the design discussion uses.  usage: summarize_ncu.py REPORT.ncu-rep OUT.csv"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1_ld_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1_ld_requests"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s [%s]" % (short, units[ix[m]]) if m in ix else short for m, short in WANT])
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].replace("cfrb200::", "")
            name = name.split("(")[0]
            w.writerow([name] + [r[ix[m]] if m in ix else "" for m, _ in WANT])
    print(open(out).read())


if __name__ == "__main__":
    main()
