"""Diagnostic: the tabulated Courtemanche flavour against the compiled reference header, step by step,
on the 150 x 2200 experiment of tests/test_gpu_wide_flavours.py.  Run with FIB_SMALL_CELLS=100000000 to
force the one-cell-per-thread flavour on the same grid.   python tests/diag_lut.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fib_tf_b200 import _capi  # noqa: E402
from oracle import cpu_port, monodomain_np as onp  # noqa: E402

ref = cpu_port.court_ref()
table = np.zeros([150, 30], np.float32)
ref.ref_init_table(table)
HI, WI, dt = 150, 2200, 0.1
cell0 = np.zeros(21, np.float32)
ref.ref_init_cell(cell0, 0)
f = (0.55 + 0.9 * np.arange(WI, dtype=np.float64) / WI).astype(np.float32)
init = np.empty([HI, WI, 21], np.float32)
init[:, :, 1:] = cell0[None, None, 1:] * f[None, :, None]
gates = [2, 3, 4, 6, 7, 8, 9, 10, 11, 13, 14, 15, 17, 18, 19]
init[:, :, gates] = np.clip(init[:, :, gates], 1e-4, 0.9999)
init[:, :, 0] = (np.arange(HI, dtype=np.float32) - 100.0)[:, None] + \
    (0.2 + 0.6 * (np.arange(WI) % 7) / 7.0).astype(np.float32)[None, :]
padded = np.pad(init, ((1, 1), (1, 1), (0, 0)), mode='symmetric')
H, W = padded.shape[:2]
ctx = _capi.Context(_capi.COURT_ULTRA, H, W, dt, 0.0, flags=_capi.F_LUT | _capi.F_NO_CLIP | _capi.F_NO_GRAPH)
ctx.set_table(_capi.TABLE_COURT_LUT, table)
for k, name in enumerate(ctx.var_names):
    ctx.set_state(name, padded[:, :, k])
st = np.ascontiguousarray(init.reshape(-1, 21))
inc = np.zeros_like(st)
names = list(ctx.var_names)
for step in range(1, 6):
    ctx.step(0, 1)
    ref.ref_euler_batch(st, inc, st.shape[0], 1, dt, table, 1)
    want = st.reshape(init.shape)
    got = np.stack([ctx.get_state(n) for n in names], axis=2)[1:-1, 1:-1]
    line = []
    for k in (0, 1, 2, 12, 16, 17):
        fl = onp.var_floor('court_ultra', names[k])
        err = np.abs(got[:, :, k].astype(np.float64) - want[:, :, k]) / np.maximum(np.abs(want[:, :, k]), fl)
        line.append('%s max %.1e >1e-3: %.4f' % (names[k], err.max(), (err > 1e-3).mean()))
    print('step %d [%s]  ' % (step, _capi.last_kernel()[12:50]) + ' | '.join(line), flush=True)
    if step == 2:
        k = 0
        err = np.abs(got[:, :, k].astype(np.float64) - want[:, :, k])
        y, x = np.unravel_index(np.argmax(err), err.shape)
        print('   worst V cell', y, x, 'init V', init[y, x, 0], 'ref', want[y, x, 0], 'cuda', got[y, x, 0],
              'f', f[x], 'finite ref', np.isfinite(want).all(), 'finite cuda', np.isfinite(got).all())
ctx.close()
