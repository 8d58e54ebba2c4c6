"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
min_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0      # skip launches shorter than this (resets, zero-substep launches)
seen = set()
for vals in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    dur = d.get("gpu__time_duration.sum", ("ms", "0"))
    dur_ms = float(dur[1].replace(",", "")) * {"us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}.get(dur[0], 1.0)
    if dur_ms < min_ms or (min_ms > 0 and d.get("Kernel Name", ("", "?"))[1] in seen):
        continue
    seen.add(d.get("Kernel Name", ("", "?"))[1])
    print("kernel:", d.get("Kernel Name", ("", "?"))[1], "grid", d.get("launch__grid_size", ("", "?"))[1], "block", d.get("launch__block_size", ("", "?"))[1])
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    for k in keys:
        if k in d:
            print(f"  {k:70s} {d[k][1]:>18s} {d[k][0]}")
    for h in hdr:
        if "issue_stalled" in h and "per_issue_active" in h:
            v = float(d[h][1])
            if v >= 0.02:
                print("  stall/issue", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), f"{v:.3f}")
