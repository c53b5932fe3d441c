"""Quick device-timed throughput probe (not the bench): python tests/quick_perf.py [model] [N] [iters]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fib_tf_b200.br import BeelerReuter
from fib_tf_b200.court import Courtemanche
from fib_tf_b200.court_ultra import Courtemanche as CourtUltra
from fib_tf_b200.fenton import Fenton4v

BYTES = {'4v': 32, 'br': 64, 'br_exact': 64, 'br_skip': 64, 'court': 168, 'court_ultra': 168, 'court_lut': 168}


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else '4v'
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    phase = '--phase' in sys.argv
    cfg = {'width': N, 'height': N, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.0, 'duration': 1,
           'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'skip': kind == 'br_skip',
           'cheby': kind in ('br', 'br_skip'), 'ultra_slow': False, 'lut': kind == 'court_lut',
           'graph': '--nograph' not in sys.argv}
    cls = {'4v': Fenton4v, 'br': BeelerReuter, 'br_exact': BeelerReuter, 'br_skip': BeelerReuter,
           'court': Courtemanche, 'court_ultra': CourtUltra, 'court_lut': CourtUltra}[kind]
    m = cls(cfg)
    if phase:
        m.add_hole_to_phase_field(N // 2, N // 2, N // 8)
    m.define()
    c = m._ctx
    c.step(0, 3)
    c.sync()
    c.timer_start()
    c.step(0, iters)
    c.timer_stop()
    ms = c.timer_ms()
    steps = iters * m.dt_per_step
    gcs = N * N * steps / (ms * 1e-3) / 1e9
    print('%-12s %5dx%-5d phase=%d  %8.3f ms / %d steps  %8.2f Gcell-steps/s  %7.1f GB/s alg (%.1f%% of 6551)' % (
        kind, N, N, phase, ms, steps, gcs, gcs * BYTES[kind], gcs * BYTES[kind] / 65.514))


if __name__ == '__main__':
    main()
