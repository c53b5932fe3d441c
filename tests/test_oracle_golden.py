"""CPU: the oracle (oracle/monodomain_np.py) against the golden fixtures produced by the
UNMODIFIED reference under oracle/tfshim.py.  Bar: bit-identical fp32 planes."""
import numpy as np
import pytest

from conftest import golden_names, load_fixture
from oracle import monodomain_np as onp


@pytest.mark.parametrize('name', golden_names())
def test_oracle_reproduces_reference_bitwise(name):
    meta, arr = load_fixture(name)
    if 'phase' in arr:
        m0 = onp.OracleModel(meta['model'], meta['config'])
        for h in meta['holes']:
            m0.add_hole(*h)
        assert np.array_equal(m0.phase, arr['phase'])
    seen = []

    def check(i, m):
        for v in meta['vars']:
            ref = arr['s%d__%s' % (i, v)]
            got = m.state[v]
            assert got.dtype == np.float32
            assert np.array_equal(got, ref, equal_nan=True), (
                '%s iter %d var %s: max |d| = %g' % (name, i, v, np.nanmax(np.abs(got - ref))))
        seen.append(i)

    m, trace = onp.run_fixture(meta, check)
    assert seen == meta['snaps']
    if meta['probe']:
        assert np.array_equal(trace, arr['probe'])
