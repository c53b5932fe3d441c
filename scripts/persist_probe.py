"""Small driver for profiling the persistent kernel: BASELINE configs 1 / 2 for a few iterations.
    python scripts/persist_probe.py [4v|br] [iterations]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fib_tf_b200.br import BeelerReuter  # noqa: E402
from fib_tf_b200.fenton import Fenton4v  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else '4v'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = {'width': 512, 'height': 512, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5 if kind == '4v' else 0.809,
       'duration': 1, 'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'skip': False, 'cheby': True}
m = (Fenton4v if kind == '4v' else BeelerReuter)(cfg)
m.add_hole_to_phase_field(*((256, 256, 30) if kind == '4v' else (150, 200, 40)))
m.define()
c = m._ctx
steps = iters * m.dt_per_step
for mode in ('one fib_step call per iteration (the driver loop)', 'one fib_step call for all iterations'):
    c.step(0, 3)
    c.sync()
    n0 = c.launch_count()
    c.timer_start()
    if mode.startswith('one fib_step call per'):
        for _ in range(iters):
            c.step(0, 1)
    else:
        c.step(0, iters)
    c.timer_stop()
    ms = c.timer_ms()
    print('%s 512^2, %s: %.2f us per time step, %.1f Gcell-steps/s (%d launches)' % (
        kind, mode, ms * 1e3 / steps, 512 * 512 * steps / ms / 1e6, c.launch_count() - n0))
m.close()

# the reference's headless driver loop (ionic.py:195-229), with and without a cycle-length observer
import time  # noqa: E402
for observer in (False, True):
    cfg2 = dict(cfg, duration=1000)
    m = (Fenton4v if kind == '4v' else BeelerReuter)(cfg2)
    m.add_hole_to_phase_field(*((256, 256, 30) if kind == '4v' else (150, 200, 40)))
    m.define()
    if observer:
        m.cl_observer = lambda i, cl: None
    m._ctx.step(0, 3)
    m._ctx.sync()
    t0 = time.perf_counter()
    n = 0
    for i in m.run(None, block=False):
        n += 1
    m._ctx.sync()
    dt = time.perf_counter() - t0
    print('%s 512^2, run(None) %s cl_observer: %d iterations in %.1f ms wall = %.1f Gcell-steps/s' % (
        kind, 'with' if observer else 'without', n, dt * 1e3, 512 * 512 * n * m.dt_per_step / dt / 1e9))
    m.close()
