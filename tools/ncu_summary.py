"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per kernel and the hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] [--kernel regex]"""
import csv, io, re, subprocess, sys, collections

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 25
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__inst_executed.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'smsp__inst_executed.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('==== ', r[idx['Kernel Name']][:100], ' id', r[idx['ID']])
    for w in WANT:
        if w in idx:
            print(f'  {w:85s} {r[idx[w]]:>16s} {units[idx[w]]}')
if '--nosrc' in sys.argv:
    sys.exit(0)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
# per-file tables of CUDA source lines with the metrics of the SASS attributed to them (innermost inline frame)
fname, func, hdr_, allrows = None, None, None, collections.defaultdict(list)
for line in src.splitlines():
    try:
        row = next(csv.reader([line]))
    except Exception:
        continue
    if not row: continue
    if row[0] == 'File Path': fname = row[1].split('/')[-1]; hdr_ = None; continue
    if row[0] == 'Function Name': func = row[1]; continue
    if row[0] == 'Line No': hdr_ = row; continue
    if hdr_ is None or len(row) < len(hdr_) - 2: continue
    allrows[func].append((fname, hdr_, row))
for func, rows_ in allrows.items():
    def col(h, r, name, occurrence=0):
        idxs = [i for i, n in enumerate(h) if n == name]
        return r[idxs[occurrence]] if len(idxs) > occurrence and idxs[occurrence] < len(r) else ''
    def num(x):
        try: return int(x)
        except Exception: return 0
    tot = sum(num(col(h, r, '# Samples')) for _, h, r in rows_)
    toti = sum(num(col(h, r, 'Instructions Executed')) for _, h, r in rows_)
    print('==== source hot spots:', func[:110], ' samples', tot, ' warp-instructions', toti)
    byfile = collections.Counter(); byfile_i = collections.Counter()
    for f, h, r in rows_:
        byfile[f] += num(col(h, r, '# Samples')); byfile_i[f] += num(col(h, r, 'Instructions Executed'))
    print('  per file (samples% / instructions%):', {f: (round(100.0 * byfile[f] / max(tot, 1), 1), round(100.0 * byfile_i[f] / max(toti, 1), 1)) for f in byfile})
    rows_.sort(key=lambda t: -num(col(t[1], t[2], '# Samples')))
    for f, h, r in rows_[:top]:
        n = num(col(h, r, '# Samples'))
        extra = ''
        for k in ('stall_long_sb', 'stall_lg', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_branch_resolving', 'stall_no_inst', 'stall_not_selected', 'stall_mio'):
            v = col(h, r, k)
            if v not in ('', '0'): extra += f' {k[6:]}={v}'
        print(f'  {100.0 * n / max(tot, 1):5.1f}%  inst={col(h, r, "Instructions Executed"):>9s} {f}:{r[0]:>4s} {r[1].strip()[:90]}  |{extra}')
