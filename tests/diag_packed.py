"""Diagnostic (not a test): which state variables differ, and by how many ulps, between two builds /
flavours after ONE time step from the same random state.
    python tests/diag_packed.py dump out.npz          (in the environment to examine)
    python tests/diag_packed.py cmp a.npz b.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def dump(path):
    from fib_tf_b200 import _capi
    from fib_tf_b200.br import BeelerReuter
    out = {}
    H, W = 96, 160
    rng = np.random.default_rng(0)
    cfgbr = {'width': W, 'height': H, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 0.809, 'duration': 1,
             'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'skip': False, 'cheby': True}
    table = BeelerReuter(cfgbr).chebyshev_table()
    cases = [('4v', _capi.FENTON4V, 0, 1.5), ('br_exact', _capi.BR, 0, 0.809), ('br_cheby', _capi.BR, _capi.F_CHEBY, 0.809),
             ('br_exact_skip', _capi.BR, _capi.F_SKIP, 0.809), ('br_cheby_skip', _capi.BR, _capi.F_CHEBY | _capi.F_SKIP, 0.809),
             ('court', _capi.COURT, 0, 0.809), ('court_ultra', _capi.COURT_ULTRA, 0, 1.5)]
    for tag, model, flags, diff in cases:
        for d in (0.0, diff):
            c = _capi.Context(model, H, W, 0.1, d, flags=flags | _capi.F_NO_GRAPH)
            r = np.random.default_rng(1)
            for v in c.var_names:
                if v in ('V',):
                    a = r.uniform(-85.0, 20.0, (H, W))
                elif v in ('U', 'W', 'S') or model == _capi.BR:
                    a = r.uniform(1e-3, 0.998, (H, W))
                    if v == 'C':
                        a = r.uniform(5e-5, 5e-3, (H, W))
                else:
                    a = None
                if a is not None:
                    c.set_state(v, a.astype(np.float32))
            if model in (_capi.COURT, _capi.COURT_ULTRA):
                from fib_tf_b200.court import INITIAL_STATE
                for name, val in INITIAL_STATE:
                    if name != 'V':
                        f = r.uniform(0.8, 1.2, (H, W))
                        c.set_state(name, np.clip(val * f, 1e-5 if name.startswith('_') and val < 1 else -1e9, 1e9).astype(np.float32))
            if flags & _capi.F_CHEBY:
                c.set_table(_capi.TABLE_BR_CHEBY, table)
            c.step(0, 1)
            if model == _capi.COURT:
                c.step(1, 1)
            out['%s/d%g/kernel' % (tag, d)] = np.array(_capi.last_kernel())
            for v in c.var_names:
                out['%s/d%g/%s' % (tag, d, v)] = c.get_state(v)
            c.close()
    np.savez(path, **out)


def ulps(a, b):
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return np.abs(ia - ib)


def cmp(pa, pb):
    a, b = np.load(pa), np.load(pb)
    for k in a.files:
        if k.endswith('/kernel'):
            print('%-28s %s   |   %s' % (k, a[k], b[k]))
            continue
        u = ulps(a[k], b[k])
        if u.max() > 0:
            print('   %-26s differs in %6d cells, max %d ulp' % (k, int((u > 0).sum()), int(u.max())))


if __name__ == '__main__':
    if sys.argv[1] == 'dump':
        dump(sys.argv[2])
    else:
        cmp(sys.argv[2], sys.argv[3])
