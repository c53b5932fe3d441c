"""How far the UNMODIFIED reference moves in cycle length / APD when only its math library changes
(tests/golden/spiral_<m>.npz vs spiral_<m>_alt.npz, both made by oracle/make_golden_spiral.py): the floor for
the statistical bounds of tests/test_spiral_statistics.py.   python scripts/spiral_spread.py >> profiles/r2_spiral_report.txt"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_spiral_statistics import apds, events  # noqa: E402

for which in ('fenton', 'br', 'court', 'court_ultra'):
    a = os.path.join(ROOT, 'tests', 'golden', 'spiral_%s.npz' % which)
    b = os.path.join(ROOT, 'tests', 'golden', 'spiral_%s_alt.npz' % which)
    if not (os.path.exists(a) and os.path.exists(b)):
        continue
    za, zb = np.load(a), np.load(b)
    meta = json.loads(str(za['meta']))
    lo, hi = {'fenton': (0.0, 1.0), 'br': (-90.0, 30.0)}.get(which, (-100.0, 50.0))
    level = lo + 0.5 * (hi - lo)
    dt_iter = meta['dt_per_step'] * meta['config']['dt']
    s2_ms = meta['s2_iter'] * dt_iter
    worst_cl = worst_apd = worst_t = 0.0
    for k in range(za['probes'].shape[1]):
        up_r, dn_r = events(za['probes'][:, k], level, dt_iter)
        up_c, dn_c = events(zb['probes'][:, k], level, dt_iter)
        post_r, post_c = up_r[up_r > s2_ms + 100], up_c[up_c > s2_ms + 100]
        if len(post_r) >= 3 and len(post_c) >= 3:
            cl_r, cl_c = np.diff(post_r).mean(), np.diff(post_c).mean()
            a_r, a_c = apds(post_r, dn_r).mean(), apds(post_c, dn_c).mean()
            worst_cl = max(worst_cl, abs(cl_c - cl_r) / cl_r)
            worst_apd = max(worst_apd, abs(a_c - a_r) / a_r)
            print('%s probe %d: reference vs reference(ALT_LIBM): cycle length %.3f vs %.3f ms (%.3f %%), APD %.3f vs %.3f ms '
                  '(%.3f %%), %d / %d beats' % (which, k, cl_c, cl_r, 100 * abs(cl_c - cl_r) / cl_r, a_c, a_r,
                                                100 * abs(a_c - a_r) / a_r, len(post_c), len(post_r)))
        elif len(up_r) == len(up_c) and len(up_r) >= 2:
            ar, ac = apds(up_r, dn_r), apds(up_c, dn_c)
            n = min(len(ar), len(ac))
            worst_t = max(worst_t, float(np.max(np.abs(up_c - up_r))))
            if n:
                worst_apd = max(worst_apd, float(np.max(np.abs(ac[:n] - ar[:n]) / ar[:n])))
            print('%s probe %d (beat by beat): reference vs reference(ALT_LIBM): activation times differ by <= %.3f ms, APD by '
                  '<= %.3f %%, %d beats' % (which, k, np.max(np.abs(up_c - up_r)),
                                            100 * np.max(np.abs(ac[:n] - ar[:n]) / ar[:n]) if n else 0.0, len(up_r)))
        else:
            print('%s probe %d: beat counts differ (%d vs %d)' % (which, k, len(up_c), len(up_r)))
    print('%s: the reference moves itself by up to %.3f %% in cycle length, %.3f %% in APD, %.3f ms in activation time'
          % (which, 100 * worst_cl, 100 * worst_apd, worst_t))
