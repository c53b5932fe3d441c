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


@pytest.mark.parametrize('which,kind,iters', [('fenton', 'fenton4v', 4), ('br', 'br', 6),
                                             ('court', 'court', 21), ('court_ultra', 'court_ultra', 12)])
def test_oracle_reproduces_the_start_of_the_512_driver_runs_bitwise(which, kind, iters):
    """The first iterations of the 512^2 driver runs stored by oracle/make_golden_spiral.py (the
    unmodified reference) against the oracle: bit-identical probe values."""
    import json
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, 'spiral_%s.npz' % which))
    meta = json.loads(str(z['meta']))
    m = onp.OracleModel(kind, meta['config'])
    m.add_hole(*meta['hole'])
    if meta.get('extra_hole'):
        m.add_hole(*meta['extra_hole'])
    m.define()
    for i in range(iters):
        m.iterate()
        if kind.startswith('court') and i % 10 == 0:
            m.fire('slow')
        got = np.array([m.pot()[r, c] for r, c in meta['probes']], np.float32)
        assert np.array_equal(got, z['probes'][i]), (which, i)


def test_waiver_lists_are_the_rule_applied_to_the_fixtures():
    """oracle.monodomain_np.WAIVERS names exactly the variables whose OWN uncertainty in the reference
    (max of the libm-swap noise and the float64 distance recorded in the fixtures) reaches 0.7e-5 in
    some fixture of the flavour; everything else is held to the flat 1e-5 bar."""
    own = {}
    for name in golden_names():
        meta, _ = load_fixture(name)
        if not meta.get('noise'):
            continue
        fl = onp.flavour_of(meta['model'], meta['config'])
        for key, v in meta['noise'].items():
            var = key.split('__', 1)[1]
            d = own.setdefault(fl, {})
            d[var] = max(d.get(var, 0.0), v, meta['rounding'].get(key, 0.0))
    derived = {fl: tuple(sorted(v for v, e in d.items() if e >= onp.WAIVE_FROM)) for fl, d in own.items()}
    derived = {fl: v for fl, v in derived.items() if v}
    assert derived == {fl: tuple(sorted(w['vars'])) for fl, w in onp.WAIVERS.items()}
    for fl, w in onp.WAIVERS.items():          # and the caps cover 3x the recorded uncertainty
        assert all(3 * own[fl][v] <= w['cap'] * 1.0001 for v in w['vars']), fl
