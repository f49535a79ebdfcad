"""Condense an .ncu-rep (read here with `ncu -i`) into a small per-launch CSV for profiles/."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct"]
idx = [(w, h.index(w)) for w in want if w in h]
with open(out, "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow([w for w, _ in idx])
    wr.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        wr.writerow([r[i].replace("void <unnamed>::", "").split("(")[0] if w == "Kernel Name" else r[i] for w, i in idx])
print(open(out).read())
