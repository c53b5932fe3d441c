"""python tests/gpu_parity_report.py  -- prints, for every golden fixture, the worst error of the
CUDA path against the fixture (the unmodified reference under tfshim) per snapshot/variable."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import golden_names, load_fixture  # noqa: E402
from cuda_adapter import CudaModel, rel_err, var_floor  # noqa: E402
from oracle import monodomain_np as onp  # noqa: E402


def main():
    worst_all = 0.0
    for name in golden_names():
        meta, arr = load_fixture(name)
        rows = []

        def check(i, m):
            for v in meta['vars']:
                ref = arr['s%d__%s' % (i, v)]
                got = m.state[v]
                rows.append((rel_err(got, ref, var_floor(meta['model'], v)), i, v,
                             float(np.nanmax(np.abs(got - ref)))))

        m, trace = onp.run_fixture(meta, check, model_factory=CudaModel)
        m.close()
        rows.sort(reverse=True)
        w = rows[0]
        worst_all = max(worst_all, w[0])
        extra = ''
        if meta['probe']:
            extra = ' probe max|d|=%.3g' % float(np.max(np.abs(trace - arr['probe'])))
        print('%-20s worst rel %.3e (iter %d var %s, abs %.3e)%s' % (name, w[0], w[1], w[2], w[3], extra))
        for r in rows[1:4]:
            print('%-20s       rel %.3e (iter %d var %s, abs %.3e)' % ('', r[0], r[1], r[2], r[3]))
    print('WORST %.3e' % worst_all)


if __name__ == '__main__':
    main()
