"""Upload + n iterations + sync, wall clock, with the pipelined upload on / off.
    python scripts/pipeline_probe.py [N] [iterations]      (env FIB_PIPELINE_MIN_CELLS / FIB_PIPELINE_BLOCK_ROWS)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fib_tf_b200 import _capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
c = _capi.Context(_capi.FENTON4V, N, N, 0.1, 1.5, steps_per_launch=2)
strip = _capi.pinned_empty((512, N))
strip[...] = 0.3
for rep in range(3):
    c.sync()
    t0 = time.perf_counter()
    for g in range(0, N, 512):
        for v in c.var_names:
            c.set_rect_async(v, g, 0, strip)
    t1 = time.perf_counter()
    for i in range(iters):
        c.step(0, 1)
    t2 = time.perf_counter()
    c.flush()
    t3 = time.perf_counter()
    c.sync()
    t4 = time.perf_counter()
    print('N=%d %d iterations: enqueue uploads %.1f ms, step calls %.1f ms, flush %.1f ms, sync %.1f ms, total %.1f ms'
          % (N, iters, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))
# upload alone, compute alone
c.sync()
t0 = time.perf_counter()
for g in range(0, N, 512):
    for v in c.var_names:
        c.set_rect_async(v, g, 0, strip)
c.sync()
t1 = time.perf_counter()
c.step(0, iters)
c.sync()
t2 = time.perf_counter()
print('upload alone %.1f ms (%.1f GB/s), %d iterations alone %.1f ms' % (
    (t1 - t0) * 1e3, 4 * N * N * 4 / (t1 - t0) / 1e9, iters, (t2 - t1) * 1e3))
c.close()
