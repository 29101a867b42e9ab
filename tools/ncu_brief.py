"""Brief of one kernel in an ncu report: python tools/ncu_brief.py rep.ncu-rep kernel_regex [--src]"""
import csv, re, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
r = [x for x in rows[2:] if re.search(pat, x[h.index('Kernel Name')])][0]
d = dict(zip(h, r))
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum']
for k in keys:
    print(k, '=', d.get(k))
for k in sorted(d):
    if 'issue_stalled' in k and 'per_issue_active' in k:
        try:
            v = float(d[k])
        except ValueError:
            continue
        if v > 0.05:
            print(k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''), round(v, 3))
if '--src' in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]; isrc = hdr.index('Source'); ie = hdr.index('Instructions Executed'); iss = hdr.index('# Samples')
    data = [(r[isrc].strip(), int(r[ie]), int(r[iss])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    first = data[0][0]
    idxs = [i for i, x in enumerate(data) if x[0] == first and i > 50]
    if idxs: data = data[:idxs[0]]
    tot = sum(x[1] for x in data); smp = sum(x[2] for x in data)
    print('sass lines', len(data), 'warp inst', tot, 'samples', smp)
    segs = []; start = 0
    for i in range(1, len(data) + 1):
        if i == len(data) or abs(data[i][1] - data[start][1]) > 0.05 * max(data[start][1], 1):
            segs.append((start, i, data[start][1])); start = i
    for s, e, c in segs:
        w = sum(data[i][1] for i in range(s, e)); sm = sum(data[i][2] for i in range(s, e))
        if w > 0.006 * tot or sm > 0.01 * smp:
            fp = sum(1 for i in range(s, e) if any(x in data[i][0] for x in ('DFMA', 'DADD', 'DMUL')))
            print(f"[{s:4d},{e:4d}) n={e-s:3d} exec={c/1e6:7.3f}M total={w/1e6:6.1f}M {w/tot*100:4.1f}% fp64={fp:3d} smp={sm:5d} {sm/smp*100:4.1f}%  {data[s][0][:44]}")
