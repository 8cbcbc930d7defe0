"""Condense an .ncu-rep (ncu --set full) into the handful of counters the design discussion uses.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.md"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    print(f"# ncu summary of `{path.split('/')[-1]}` (ncu --set full --clock-control none; per launch)\n")
    print("| metric | unit | " + " | ".join(r[ki].split("(")[0] for r in data) + " |")
    print("|---|---|" + "---|" * len(data))
    for key, label in WANT:
        if key not in H:
            continue
        i = H.index(key)
        vals = []
        for r in data:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.4g}" if abs(v) < 1e6 else f"{v:.4e}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} (`{key}`) | {U[i]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
