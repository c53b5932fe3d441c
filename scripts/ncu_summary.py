"""Summarise gpurun_out/*.ncu-rep + launch lists into profiles/ (text + traffic.json).
   python scripts/ncu_summary.py r1"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
G = os.path.join(ROOT, 'gpurun_out')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
traffic = {}
lines = []
# (report, key in traffic.json, algorithmic bytes per cell-step, time steps per launch)
# (the persistent kernel: one launch = 64 iterations of a 512^2 grid, state on chip in between)
for rep, key, balg, spl, CELLS in (
        ('prof_4v', 'fenton4v_step', 32, 1, 4096 * 4096), ('prof_4v_fused', 'fenton4v_fused2_step', 32, 2, 4096 * 4096),
        ('prof_br', 'br_cheby_step', 64, 1, 4096 * 4096), ('prof_br_exact', 'br_exact_step', 64, 1, 4096 * 4096),
        ('prof_court', 'court_ultra_step', 168, 1, 4096 * 4096),
        ('prof_court_lut', 'court_lut_step', 168, 1, 4096 * 4096),
        ('prof_persist_4v', 'fenton4v_persist_512', 32, 640, 512 * 512),
        ('prof_persist_br', 'br_cheby_persist_512', 64, 320, 512 * 512)):
    path = os.path.join(G, rep + '.ncu-rep')
    pre = os.path.join(G, rep + '.raw.csv')          # exported on the GPU box (scripts/gpu_ncu.sh)
    if os.path.exists(pre) and os.path.getsize(pre) > 0:
        txt = open(pre).read()
    elif os.path.exists(path):
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
    side = int(round(CELLS ** 0.5))
    lines.append('== %s : %s   (%dx%d grid, one launch = %s, ncu --set full --clock-control none)'
                 % (rep, d['Kernel Name'][0], side, side, 'one time step' if spl == 1 else '%d time steps' % spl))
    for k in KEEP:
        if k in d:
            lines.append('   %-86s %s %s' % (k, d[k][0], d[k][1]))

    def num(k):
        v, u = d[k]
        v = float(v.replace(',', ''))
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9,
                    'inst': 1}.get(u, 1)
    # per CELL-STEP: a launch of the fused kernel advances every cell by `spl` steps
    dram = (num('dram__bytes_read.sum') + num('dram__bytes_write.sum')) / spl
    inst = num('smsp__inst_executed.sum') * 32 / CELLS / spl
    t = num('gpu__time_duration.sum') / spl
    traffic[key] = {'dram_bytes_per_cell': dram / CELLS, 'algorithmic_bytes_per_cell': balg,
                    'thread_instructions_per_cell': inst, 'ncu_duration_us': t * 1e6, 'time_steps_per_launch': spl,
                    'gcell_steps_per_s_under_ncu': CELLS / t / 1e9, 'source': 'profiles/%s_ncu_summary.txt' % tag}
    lines.append('   -> DRAM bytes / cell-step %.1f (algorithmic %d), thread-instructions / cell-step %.0f, '
                 '%.1f Gcell-steps/s under ncu (cold cache, serialised)' % (dram / CELLS, balg, inst, CELLS / t / 1e9))
    lines.append('')
# launch list of the bench command
lp = os.path.join(G, 'launches_bench4096.csv')
if os.path.exists(lp):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(lp)):
        if len(r) > 10 and r[0].isdigit():
            agg[r[4]][0] += 1
            agg[r[4]][1] += float(r[-1])
    tot = sum(v[1] for v in agg.values())
    lines.append('== launch list of `python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu --no-suite` '
                 '(ncu --metrics gpu__time_duration.sum, every launch of the run: the persistent kernel building the '
                 '512^2 spiral tile, the warm-up, the timed region, the e2e loop; shares, not absolutes)')
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append('   %5.1f %%  %4d launches  %9.1f us avg   %s' % (100 * ns / tot, n, ns / n / 1e3, k[:90]))
open(os.path.join(OUT, '%s_ncu_summary.txt' % tag), 'w').write('\n'.join(lines) + '\n')
json.dump(traffic, open(os.path.join(OUT, 'traffic.json'), 'w'), indent=1)
print('\n'.join(lines))
