"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/X.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max", "lts__t_bytes.sum", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    lines = [f"# ncu --set full --clock-control none summary of {rep} (values per captured launch)"]
    for d in data:
        lines.append(f"kernel: {d[name_i]}")
    for i, h in enumerate(hdr):
        if h in WANT:
            lines.append(f"{h} [{units[i]}]: " + ", ".join(d[i] for d in data))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
