"""(CPU) profiles/r02_traffic.json from the round's `ncu --set full` reports: per kernel the DRAM bytes of the profiled launch,
its duration under ncu, issue-slot utilisation, lanes per instruction, occupancy, cache hit rates, registers, the two stall
figures the docs quote.  bench.py reads it for `roofline.traffic` / `issue_active_pct` / `lanes_per_inst`.
    python tools/ncu_traffic.py gpurun_out/r02_cbox.ncu-rep:k_extend,k_shade gpurun_out/r02_c4.ncu-rep:k_extend_sm,k_shadow_sm ..."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for arg in sys.argv[1:]:
    rep, names = arg.split(':'); names = names.split(',')
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt))); hdr, units = rows[0], rows[1]

    def col(r, name):
        i = hdr.index(name)
        return float(r[i].replace(',', '')), units[i]
    for r, k in zip(rows[2:], names):
        size = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        time = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0}
        rd, u1 = col(r, 'dram__bytes_read.sum'); wr, u2 = col(r, 'dram__bytes_write.sum')
        b = rd * size[u1] + wr * size[u2]
        t, ut = col(r, 'gpu__time_duration.sum'); t *= time[ut]
        out[k] = {'kernel': r[hdr.index('Kernel Name')], 'dram_bytes_per_launch': b, 'launch_us_under_ncu': t * 1e6, 'dram_gbs': b / t / 1e9,
                  'issue_active_pct': col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')[0],
                  'lanes_per_inst': col(r, 'smsp__thread_inst_executed_per_inst_executed.ratio')[0],
                  'warps_active_pct': col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')[0],
                  'l2_hit_pct': col(r, 'lts__t_sector_hit_rate.pct')[0], 'l1_hit_pct': col(r, 'l1tex__t_sector_hit_rate.pct')[0],
                  'registers': col(r, 'launch__registers_per_thread')[0],
                  'long_scoreboard_per_issue': col(r, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio')[0],
                  'no_instruction_per_issue': col(r, 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio')[0],
                  'report': os.path.basename(rep)}
json.dump(out, open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w'), indent=1)
for k, v in out.items():
    print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a not in ('kernel', 'report')})
