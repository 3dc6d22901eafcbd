"""Selected metrics of .ncu-rep files as one CSV (kernel, metric, value, unit):  python tools/ncu_select.py out.csv rep [rep...]"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__inst_executed_pipe_xu.sum",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
out = csv.writer(open(sys.argv[1], "w"))
out.writerow(["report", "kernel", "metric", "value", "unit"])
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        for w in WANT:
            if w in h:
                i = h.index(w)
                out.writerow([rep.split("/")[-1], r[ki][:70], w, r[i], units[i]])
