"""GPU: the persistent on-chip kernel (csrc/fib_persist.cuh) -- up to 64 run() iterations per launch
for small unsharded Fenton 4v / Beeler-Reuter grids (the reference's own 512^2 configurations), state
resident in registers / shared memory between the time steps, TMA tile loads and stores, neighbour
tiles exchanging their edge rows through a mailbox of self-validating {value, step number} words.

It must reproduce one launch per time step BIT FOR BIT (same cell functions, same clamped index map,
-fmad=false build): every model flavour, with and without a phase field, tile heights 2 / 4 / 8,
widths that are not a multiple of the TMA box or of 4, partial last tiles, stimuli between
iterations, and the reference's 512^2 configurations with their holes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cuda(cuda_device):
    import cuda_adapter
    return cuda_adapter


def _pair(model, H, W, dt, diff, flags, init, phase, table):
    from fib_tf_b200 import _capi
    out = []
    for extra in (0, _capi.F_NO_PERSIST):
        c = _capi.Context(model, H, W, dt, diff, flags=flags | extra)
        for v, a in init.items():
            c.set_state(v, a)
        if phase is not None:
            c.set_phase(phase, 0)
        if table is not None:
            c.set_table(_capi.TABLE_BR_CHEBY, table)
        out.append(c)
    return out


def _phase(H, W):
    yy, xx = np.mgrid[0:H, 0:W]
    ph = np.ones((H, W))
    for cy, cx, rad in ((H / 2.0, W / 2.0, max(min(H, W) / 6.0, 1.0)), (0.0, W * 0.8, max(min(H, W) / 8.0, 1.0))):
        ph *= 0.5 * (np.tanh(np.hypot(yy - cy, xx - cx) - rad) + 1.0)
    return np.maximum(ph, 1e-5).astype(np.float32)


@pytest.mark.parametrize('hole', [False, True])
@pytest.mark.parametrize('H,W,th', [(3, 3, 2), (7, 5, 2), (64, 97, 2), (300, 256, 4), (301, 500, 4), (512, 512, 4),
                                    (1000, 512, 8), (1184, 300, 8)])
def test_fenton_persistent_kernel_is_bit_identical(cuda, H, W, th, hole):
    from fib_tf_b200 import _capi
    rng = np.random.default_rng(H * 7 + W)
    init = {v: rng.uniform(0.0, 1.0, (H, W)).astype(np.float32) for v in ('U', 'V', 'W', 'S')}
    init['U'][H // 3:H // 2 + 1, W // 4:W // 2] = 0.95
    per, ref = _pair(_capi.FENTON4V, H, W, 0.1, 1.5, 0, init, _phase(H, W) if hole else None, None)
    stim = ('U', 1, max(H - 1, 2), 1, max(W // 2, 2), 0.6, 0.0)
    n_per, n_ref = per.launch_count(), ref.launch_count()
    for it in range(4):
        per.step(0, 1)
        ref.step(0, 1)
        if it == 0:
            k = _capi.last_kernel()          # the plain context launched last
            assert 'persist' not in k, k
        if it == 1:
            per.stimulate(*stim)
            ref.stimulate(*stim)
    per.step(0, 3)
    per.flush()                              # iterations are deferred until something looks (include/fib_b200.h)
    assert 'persist_kernel<Fenton4v,TH=%d,PHASE=%d>' % (th, hole) in _capi.last_kernel(), _capi.last_kernel()
    ref.step(0, 3)
    for v in ref.var_names:
        want = ref.get_state(v)
        assert np.isfinite(want).all()
        assert np.array_equal(per.get_state(v), want), v
    # deferred iterations: two before the stimulus in one launch, the five after it in another (+ the stimulus)
    assert per.launch_count() - n_per == 2 + 1 and ref.launch_count() - n_ref == 70 + 1
    per.close()
    ref.close()


@pytest.mark.parametrize('flags', ['exact', 'exact+skip', 'cheby', 'cheby+skip', 'cheby+skip+strict'])
@pytest.mark.parametrize('H,W,th,hole', [(5, 9, 2, True), (130, 244, 2, False), (512, 512, 4, True), (592, 160, 4, False)])
def test_beeler_reuter_persistent_kernel_is_bit_identical(cuda, H, W, th, hole, flags):
    from fib_tf_b200 import _capi
    from fib_tf_b200.br import BeelerReuter
    fl = (_capi.F_CHEBY if 'cheby' in flags else 0) | (_capi.F_SKIP if 'skip' in flags else 0) | \
        (_capi.F_CHEBY_STRICT if 'strict' in flags else 0)
    cfg = {'width': W, 'height': H, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 0.809, 'duration': 1,
           'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'skip': False, 'cheby': True}
    table = BeelerReuter(cfg).chebyshev_table() if 'cheby' in flags else None
    rng = np.random.default_rng(H * 3 + W)
    init = {'V': rng.uniform(-85.0, 20.0, (H, W)).astype(np.float32),
            'C': rng.uniform(5e-5, 5e-3, (H, W)).astype(np.float32)}
    for g in ('M', 'H', 'J', 'D', 'F', 'XI'):
        init[g] = rng.uniform(1e-3, 0.998, (H, W)).astype(np.float32)
    per, ref = _pair(_capi.BR, H, W, 0.1, 0.809, fl, init, _phase(H, W) if hole else None, table)
    stim = ('V', 1, max(H // 2, 2), 1, max(W // 2, 2), 10.0, -90.0)
    n_per, n_ref = per.launch_count(), ref.launch_count()
    for it in range(5):
        per.step(0, 1)
        ref.step(0, 1)
        if it == 2:
            per.stimulate(*stim)
            ref.stimulate(*stim)
    per.step(0, 1)
    per.flush()
    name = {'exact': 'exact', 'cheby': 'cheby'}['cheby' if 'cheby' in flags else 'exact']
    if 'strict' in flags:
        name = 'strict'
    assert 'persist_kernel<BeelerReuter<%s,slow>,TH=%d,PHASE=%d>' % (name, th, hole) in _capi.last_kernel(), \
        _capi.last_kernel()
    ref.step(0, 1)
    for v in ref.var_names:
        a, b = per.get_state(v), ref.get_state(v)
        assert np.array_equal(a, b, equal_nan=True), (v, float(np.nanmax(np.abs(a - b))))
    assert per.launch_count() - n_per == 2 + 1 and ref.launch_count() - n_ref == 30 + 1
    per.close()
    ref.close()


def test_persistent_kernel_is_not_used_where_it_does_not_apply(cuda):
    """Too many rows per SM (BR: more than 4), wide grids, row shards and the two-steps-per-launch layout
    keep the one-launch-per-step path; the run() generator and the probe ring work on top of either."""
    from fib_tf_b200 import _capi
    from fib_tf_b200.fenton import Fenton4v
    for model, H, W, kw in ((_capi.BR, 700, 512, {}), (_capi.FENTON4V, 64, 640, {}),
                            (_capi.FENTON4V, 64, 64, {'steps_per_launch': 2}),
                            (_capi.COURT, 64, 64, {})):
        c = _capi.Context(model, H, W, 0.1, 1.0, **kw)
        for v in c.var_names:
            c.set_state(v, np.full((H, W), 0.5 if v != 'V' else -80.0, np.float32))
        c.step(0, 1)
        assert 'persist' not in _capi.last_kernel(), (model, H, W, _capi.last_kernel())
        c.close()
    cfg = {'width': 96, 'height': 64, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5, 'duration': 30,
           'timeline': False, 'timeline_name': 'x', 'save_graph': False}
    seen = []
    states = []
    for persist in (True, False):
        m = Fenton4v(dict(cfg, **({} if persist else {'persist': False})))
        m.define()
        m.cl_observer = lambda i, cl: seen.append((persist, i, cl))
        for i in m.run(None, block=False):
            pass
        states.append(m._State['U'].eval())
        m.close()
    assert np.array_equal(states[0], states[1])
    assert [s[1:] for s in seen if s[0]] == [s[1:] for s in seen if not s[0]] and seen


@pytest.mark.parametrize('model,watch', [('4v', 'U'), ('4v', 'W'), ('br', 'V'), ('br', 'XI')])
def test_deferred_iterations_and_the_in_kernel_probe_ring(cuda, model, watch):
    """fib_step on the persistent path only counts; the launch (up to 64 iterations, the watched cell
    recorded by the kernel itself after every iteration) happens when something looks.  The ring and the
    state must be those of one launch per step with probe_record_kernel between the iterations."""
    from fib_tf_b200 import _capi
    H, W = 200, 160
    rng = np.random.default_rng(11)
    if model == '4v':
        init = {v: rng.uniform(0.0, 1.0, (H, W)).astype(np.float32) for v in ('U', 'V', 'W', 'S')}
        per, ref = _pair(_capi.FENTON4V, H, W, 0.1, 1.5, 0, init, _phase(H, W), None)
        stim = ('U', 90, 110, 70, 90, 0.5, 0.0)
    else:
        init = {'V': rng.uniform(-85.0, 20.0, (H, W)).astype(np.float32),
                'C': rng.uniform(5e-5, 5e-3, (H, W)).astype(np.float32)}
        for g in ('M', 'H', 'J', 'D', 'F', 'XI'):
            init[g] = rng.uniform(1e-3, 0.998, (H, W)).astype(np.float32)
        per, ref = _pair(_capi.BR, H, W, 0.1, 0.809, _capi.F_SKIP, init, None, None)
        stim = ('V', 90, 110, 70, 90, 10.0, -90.0)
    for c in (per, ref):
        c.probe_watch(watch, 101, 77)            # row 101 is not the first row of its tile
    n0 = per.launch_count()
    for it in range(150):
        per.step(0, 1)
        ref.step(0, 1)
        if it == 99:
            per.stimulate(*stim)                 # covers the probe cell: updates the last record
            ref.stimulate(*stim)
    a, b = per.probe_fetch(), ref.probe_fetch()
    assert a.size == 150 and np.array_equal(a, b)
    # 64 + 36 iterations before the stimulus, 50 after it: three step launches (+ stimulus + record update)
    assert per.launch_count() - n0 <= 3 + 2
    per.step(0, 3)
    ref.step(0, 3)
    assert np.array_equal(per.probe_fetch(), ref.probe_fetch())
    # partial fetches wait for the launch that produced the values asked for, not for the whole stream
    for it in range(130):
        per.step(0, 1)
        ref.step(0, 1)
    a = np.concatenate([per.probe_fetch(64), per.probe_fetch(10), per.probe_fetch()])
    assert a.size == 130 and np.array_equal(a, ref.probe_fetch())
    for v in ref.var_names:
        assert np.array_equal(per.get_state(v), ref.get_state(v)), v
    per.close()
    ref.close()
