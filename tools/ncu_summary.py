"""Compact per-launch table from an .ncu-rep (`ncu -i rep --page raw --csv`): the metrics the roofline discussion
needs.  usage: ncu_summary.py <rep> [> profiles/xxx.md]"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe % (active)"),
    ("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "FP64 tensor ops % of peak (elapsed)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math-pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}` ({len(data)} launches)\n")
    names = [re.sub(r"\(.*$", "", r[ki]).replace("void ", "")[:60] + " grid " + r[gi] for r in data]
    print("| metric | " + " | ".join(f"#{i}" for i in range(len(data))) + " |")
    print("|---|" + "---|" * len(data))
    for m, label in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        vals = []
        for r in data:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} [{units[i]}] | " + " | ".join(vals) + " |")
    print()
    for i, n in enumerate(names):
        print(f"- #{i}: `{n}`")


if __name__ == "__main__":
    main()
