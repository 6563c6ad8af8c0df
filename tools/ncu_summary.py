"""Print the metrics that matter from an .ncu-rep (run here, no GPU):  python tools/ncu_summary.py <rep> [kernel-index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "device__attribute_clock_rate"]
for i, k in enumerate(hdr):
    if k in KEYS or ("issue_stalled" in k and "per_issue_active" in k):
        print(f"{k:92s} {vals[i]:>18s} {units[i]}")
