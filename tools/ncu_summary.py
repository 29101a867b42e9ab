"""Print the key counters of an ncu report (first kernel instance) — used to write profiles/*.txt."""
import csv, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vals = rows[2 + idx]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
pats = [r"^Kernel Name$", r"gpu__time_duration.sum$", r"launch__registers_per_thread$", r"launch__grid_size", r"launch__occupancy_limit_(registers|shared_mem|warps)",
        r"launch__shared_mem_per_block_dynamic", r"sm__cycles_elapsed.max$", r"smsp__inst_executed.sum$",
        r"smsp__issue_active.avg.pct_of_peak_sustained_active", r"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_(active|elapsed)",
        r"sm__inst_executed_pipe_fp64.(min|max).pct_of_peak_sustained_active",
        r"smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on.sum$", r"smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on.sum.per_cycle_elapsed",
        r"smsp__warps_active.avg.per_cycle_active", r"smsp__thread_inst_executed_per_inst_executed.ratio",
        r"dram__bytes_(read|write).sum$", r"lts__t_bytes.sum$", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$",
        r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio",
        r"sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu|uniform|cbu|adu|fmaheavy|fmalite).sum$", r"smsp__inst_executed_pipe_(alu|fma|fp64|lsu|xu|uniform|cbu).sum$"]
for k in sorted(d):
    if any(re.search(p, k) for p in pats):
        v, u = d[k]
        if k.startswith("smsp__average_warps_issue_stalled") and float(v) < 0.05:
            continue
        print(f"{k} = {v} {u}")
