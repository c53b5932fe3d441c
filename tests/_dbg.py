import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
from oracle import monodomain_np as onp
import cuda_adapter as cuda
H, W = 20, 66
for kind in ('court_ultra', 'court'):
    cfg = {'width': W, 'height': H, 'dt': 0.1 if kind == 'court_ultra' else 0.01, 'dt_per_plot': 10, 'diff': 0.0, 'duration': 1,
           'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'ultra_slow': kind == 'court_ultra'}
    sing = np.float32([-10.0001, 7.9, -47.13, -14.1, 3.3328, 19.9, -40.0])
    vals = []
    for s in sing:
        for k in (0, 1, -1, 2, -2, 3, -3, 10, -10, 1000, -1000):
            v = np.float32(s)
            for _ in range(abs(k)):
                v = np.nextafter(v, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
            vals.append(v)
    vals = np.float32(vals)
    rng = np.random.default_rng(7)
    V = rng.uniform(-90.0, 40.0, size=(H, W)).astype(np.float32)
    V.flat[:4 * vals.size] = np.tile(vals, 4)
    rng.shuffle(V.reshape(-1))
    ref, gpu = onp.OracleModel(kind, cfg), cuda.CudaModel(kind, cfg)
    ref.define(s1=False); gpu.define(s1=False)
    init = {k: a.copy() for k, a in ref.state.items()}
    for k in init:
        if k == 'V':
            init[k] = V
        elif k in ('_m_', '_h_', '_j_', '_oa_', '_oi_', '_ua_', '_ui_', '_xr_', '_xs_', '_d_', '_f_', '_f_Ca_', '_u_', '_v_', '_w_', '_us_'):
            init[k] = rng.uniform(1e-3, 0.998, (H, W)).astype(np.float32)
    for k, a in init.items():
        ref.state[k] = a.copy(); gpu.m._State[k].assign(a)
    for it in range(1):
        with np.errstate(all='ignore'):
            ref.iterate(); 
            if kind == 'court': ref.fire('slow')
        gpu.iterate()
        if kind == 'court': gpu.fire('slow')
    for v in ref.state:
        got, want = gpu.state[v].astype(np.float64), ref.state[v].astype(np.float64)
        err = np.abs(got - want) / np.maximum(np.abs(want), onp.var_floor(kind, v))
        err[0, :] = err[-1, :] = 0; err[:, 0] = err[:, -1] = 0
        idx = int(np.argmax(err)); r, c = divmod(idx, W)
        if err[r, c] > 1e-5:
            print('%-12s %-9s worst %.2e at V0=%.7f got %.8g want %.8g init %.8g' % (kind, v, err[r, c], V[r, c], got[r, c], want[r, c], init[v][r, c]))
    gpu.close()
