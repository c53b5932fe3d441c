"""Device-timed throughput of every BASELINE.json configuration (plus the 4096^2 roofline points)
through the drop-in API; writes profiles/<tag>_suite.json.   python scripts/suite.py r1"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fib_tf_b200.br import BeelerReuter  # noqa: E402
from fib_tf_b200.court import Courtemanche  # noqa: E402
from fib_tf_b200.court_ultra import Courtemanche as CourtUltra  # noqa: E402
from fib_tf_b200.fenton import Fenton4v  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
CASES = [
    # name, class, N, extra config, hole (x, y, r) or None, bytes/cell-step, iterations, slow_every
    ('config1: 4v 512^2, hole(256,256,30), diff 1.5', Fenton4v, 512, {'diff': 1.5}, (256, 256, 30), 36, 300, 0),
    ('config2: BR 512^2 cheby, hole(150,200,40), diff 0.809', BeelerReuter, 512,
     {'diff': 0.809, 'cheby': True, 'skip': False}, (150, 200, 40), 68, 300, 0),
    ('config3: BR 2048^2 cheby+skip', BeelerReuter, 2048, {'diff': 0.809, 'cheby': True, 'skip': True}, None, 64, 40, 0),
    ('config4a: Courtemanche 2048^2 multi-rate (court.py loop, slow every 10)', Courtemanche, 2048,
     {'diff': 0.809}, None, 168, 100, 10),
    ('config4b: Courtemanche 2048^2 ultra (all states every step)', CourtUltra, 2048, {'diff': 1.5}, None, 168, 40, 0),
    ('config4c: Courtemanche 2048^2 ultra + LUT', CourtUltra, 2048, {'diff': 1.5, 'lut': True}, None, 168, 40, 0),
    ('roofline point: 4v 4096^2', Fenton4v, 4096, {'diff': 1.5}, None, 32, 10, 0),
    ('roofline point: 4v 4096^2 + hole', Fenton4v, 4096, {'diff': 1.5}, (2048, 2048, 240), 36, 10, 0),
    ('roofline point: BR 4096^2 cheby', BeelerReuter, 4096, {'diff': 0.809, 'cheby': True, 'skip': False}, None, 64, 10, 0),
    ('roofline point: BR 4096^2 exact gates', BeelerReuter, 4096, {'diff': 0.809, 'cheby': False, 'skip': False}, None, 64, 10, 0),
    ('roofline point: BR 4096^2 cheby+skip', BeelerReuter, 4096, {'diff': 0.809, 'cheby': True, 'skip': True}, None, 64, 10, 0),
    ('roofline point: Courtemanche ultra 4096^2', CourtUltra, 4096, {'diff': 1.5}, None, 168, 10, 0),
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    ncases = int(sys.argv[2]) if len(sys.argv) > 2 else len(CASES)
    out = []
    for name, cls, N, extra, hole, balg, iters, slow_every in CASES[:ncases]:
        cfg = {'width': N, 'height': N, 'dt': 0.1, 'dt_per_plot': 10, 'duration': 1, 'timeline': False,
               'timeline_name': 'x', 'save_graph': False, 'skip': False, 'cheby': False, 'ultra_slow': False}
        cfg.update(extra)
        m = cls(cfg)
        if hole:
            m.add_hole_to_phase_field(*hole)
        m.define()
        c = m._ctx

        def go(n):
            for i in range(n):
                c.step(0, 1)
                if slow_every and i % slow_every == 0:
                    c.step(1, 1)
        go(3)
        c.sync()
        c.timer_start()
        go(iters)
        c.timer_stop()
        ms = c.timer_ms()
        steps = iters * m.dt_per_step
        g = N * N * steps / (ms * 1e-3) / 1e9
        row = {'case': name, 'grid': N, 'time_steps': steps, 'ms': ms, 'us_per_step': ms * 1e3 / steps,
               'gcell_steps_per_s': g, 'bytes_per_cell_step': balg, 'algorithmic_GBps': g * balg,
               'frac_of_measured_hbm': g * balg / PEAK,
               'state_bytes': N * N * 4 * c.nvars, 'l2_resident': N * N * 4 * c.nvars < 100e6}
        out.append(row)
        print('%-72s %8.2f Gcell-steps/s  %6.2f us/step  %5.1f %% of HBM roofline%s' % (
            name, g, row['us_per_step'], 100 * row['frac_of_measured_hbm'],
            '  (L2-resident)' if row['l2_resident'] else ''), flush=True)
        m.close()
    path = os.path.join(ROOT, 'gpurun_out', '%s_suite.json' % tag)
    json.dump({'peak_hbm_gbs': PEAK, 'timing': 'CUDA events on the library stream, graph replay, 3 warm-up iterations',
               'cases': out}, open(path, 'w'), indent=1)


if __name__ == '__main__':
    main()
