"""CPU: the host-side driver logic of the drop-in IonicModel (run() generator ordering, samples
arithmetic, fire_op -> rectangle, 'slow' op dispatch, cycle-length observer, keep_state, timeline)
exercised against a RECORDING test double of the C-ABI context.  The double computes nothing --
the product has no CPU path -- it only records the calls the host logic makes."""
import json

import numpy as np
import pytest

from fib_tf_b200 import _capi
from fib_tf_b200.br import BeelerReuter
from fib_tf_b200.court import Courtemanche
from fib_tf_b200.fenton import Fenton4v


class RecordingContext:
    NAMES = {_capi.FENTON4V: ['U', 'V', 'W', 'S'],
             _capi.BR: ['V', 'C', 'M', 'H', 'J', 'D', 'F', 'XI'],
             _capi.COURT: ['V', '_Na_i_', '_m_', '_h_', '_j_', '_K_i_', '_oa_', '_oi_', '_ua_', '_ui_',
                           '_xr_', '_xs_', '_Ca_i_', '_d_', '_f_', '_f_Ca_', '_Ca_rel_', '_u_', '_v_',
                           '_w_', '_Ca_up_']}
    STEPS = {_capi.FENTON4V: 10, _capi.BR: 5, _capi.COURT: 1}
    probe_script = None

    def __init__(self, model, height, width, dt, diff, flags=0, device=0, row0=0, rows=0,
                 steps_per_launch=0):
        self.model, self.flags, self.height, self.width = model, flags, height, width
        self.rows, self.row0 = rows or height, row0
        self.var_names = list(self.NAMES[model])
        self.nvars, self.dt_per_step = len(self.var_names), self.STEPS[model]
        self.state, self.calls, self.launches = {}, [], 0

    def set_state(self, var, host):
        self.state[var] = np.array(host, dtype=np.float32)

    def get_state(self, var, out=None):
        return self.state[var].copy()

    def set_phase(self, rows, first_row=0):
        self.calls.append(('phase', np.asarray(rows).shape, first_row))

    def set_table(self, table, data):
        self.calls.append(('table', table, np.asarray(data).shape))

    def step(self, op=0, n_iter=1):
        self.calls.append(('step', op, n_iter))
        self.launches += self.dt_per_step * n_iter
        if self.watch and op == 0:              # the device records the watched cell per iteration
            for _ in range(n_iter):
                var, row, col = self.watch
                self.ring.append(self.probe_script.pop(0) if self.probe_script else self.state[var][row, col])

    watch, ring = None, None

    def probe_watch(self, var, row, col):
        self.calls.append(('watch', var, row, col))
        self.watch, self.ring = ((var, row, col) if row >= 0 else None), []

    def probe_fetch(self, max_values=4096):
        self.calls.append(('fetch', max_values))
        vals = self.ring[:max_values]
        del self.ring[:len(vals)]
        return np.asarray(vals, dtype=np.float32)

    def stimulate(self, var, r0, r1, c0, c1, value, floor_v):
        self.calls.append(('stim', var, r0, r1, c0, c1, value, floor_v))

    def probe(self, var, row, col):
        self.calls.append(('probe', var, row, col))
        return np.float32(self.probe_script.pop(0) if self.probe_script else self.state[var][row, col])

    def sync(self):
        self.calls.append(('sync',))

    def timer_start(self):
        pass

    def timer_stop(self):
        pass

    def timer_ms(self):
        return 0.25

    def launch_count(self):
        return self.launches

    def close(self):
        pass


@pytest.fixture
def recording(monkeypatch):
    monkeypatch.setattr(_capi, 'Context', RecordingContext)
    RecordingContext.probe_script = None
    return RecordingContext


CFG = {'width': 64, 'height': 48, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5, 'duration': 7,
       'timeline': False, 'timeline_name': 'unused.json', 'save_graph': False, 'skip': True, 'cheby': True}


def test_define_uploads_the_reference_initial_state(recording):
    m = Fenton4v(CFG)
    m.add_hole_to_phase_field(30, 20, 5)
    m.define()
    c = m._ctx
    assert m.dt_per_step == 10 and m.defined
    assert np.all(c.state['U'][:, 1] == 1.0) and c.state['U'].sum() == 48      # S1 = column 1
    assert np.all(c.state['V'] == 1.0) and np.all(c.state['S'] == 0.0)
    assert ('phase', (48, 64), 0) in c.calls
    b = BeelerReuter(CFG)
    b.define(s1=False)
    assert np.all(b._ctx.state['V'] == np.float32(-84.624)) and b._ctx.flags == (_capi.F_CHEBY | _capi.F_SKIP)
    assert ('table', _capi.TABLE_BR_CHEBY, (12, 9)) in b._ctx.calls


def test_run_generator_order_and_stimulus_rectangle(recording):
    m = Fenton4v(CFG)
    m.define()
    m.add_pace_op('s2', 'llq', 0.9)
    seen = []
    for i in m.run(None):
        seen.append((i, len([c for c in m._ctx.calls if c[0] == 'step'])))   # step i ran BEFORE yield i
        if i == 3:
            m.fire_op('s2')
    assert m.samples == 7 and [s[0] for s in seen] == list(range(7))
    assert [s[1] for s in seen] == [1, 2, 3, 4, 5, 6, 7]
    stim = [c for c in m._ctx.calls if c[0] == 'stim']
    assert stim == [('stim', 'U', 24, 47, 1, 32, 0.9, 0.0)]                  # llq: rows H//2:-1, cols 1:W//2
    assert m._ctx.calls[-1] == ('sync',)


def test_cycle_length_observer_without_a_screen(recording):
    m = BeelerReuter(dict(CFG, duration=30, dt_per_plot=5, probe_batch=16))     # probe every iteration (5/5)
    m.define()
    hits = []
    m.cl_observer = lambda i, cl: hits.append((i, cl))
    # V at [20, W//2]: below -30 mV (image < 0.5), up-crossing at iteration 4, down, up again at 40
    RecordingContext.probe_script = [-80.0] * 4 + [0.0] * 10 + [-80.0] * 26 + [10.0] * 20
    for _ in m.run(None):
        pass
    assert [h[0] for h in hits] == [4, 40]
    assert hits[1][1] == pytest.approx((40 - 4) * 5 * 0.1)        # cycle length in ms (ionic.py:218)
    # the probe cell [20, W//2] is watched on the device and read back 16 iterations at a time, one batch
    # behind the stepping (at iterations 31 and 47), the rest at the end
    calls = m._ctx.calls
    assert ('watch', 'V', 20, 32) in calls and not [c for c in calls if c[0] == 'probe']
    assert [c[1] for c in calls if c[0] == 'fetch'] == [16, 16, 28]
    assert calls[-2] == ('watch', 'V', -1, -1)
    # probe_batch = 1: read back after every iteration, same observer calls
    m = BeelerReuter(dict(CFG, duration=30, dt_per_plot=5, probe_batch=1))
    m.define()
    hits2 = []
    m.cl_observer = lambda i, cl: hits2.append((i, cl))
    RecordingContext.probe_script = [-80.0] * 4 + [0.0] * 10 + [-80.0] * 26 + [10.0] * 20
    for _ in m.run(None):
        pass
    assert hits2 == hits and len([c for c in m._ctx.calls if c[0] == 'fetch']) == 60


def test_courtemanche_slow_op_state_dict_and_timeline(recording, tmp_path):
    cfg = dict(CFG, duration=1.2, timeline=True, timeline_name=str(tmp_path / 't.json'))
    m = Courtemanche(cfg)
    m.define()
    for i in m.run(None, keep_state=True):
        if i % 10 == 0:
            m.fire_op('slow')
    ops = [c[1] for c in m._ctx.calls if c[0] == 'step']
    # int(1.2 / (1 * 0.1)) == 11 in floating point, exactly as in the reference (ionic.py:198)
    assert m.samples == 11
    assert ops.count(_capi.OP_SLOW) == 2 and ops.count(_capi.OP_ODE) == 11 + 1    # +1 traced iteration
    assert ops[:3] == [_capi.OP_ODE, _capi.OP_SLOW, _capi.OP_ODE]                # slow AFTER the fast op
    assert sorted(m.state) == sorted(RecordingContext.NAMES[_capi.COURT])
    assert m.state['V'].shape == (48, 64) and np.all(m.state['V'][:, :25] == 20.0)
    assert json.load(open(cfg['timeline_name']))['traceEvents'][0]['dur'] == 250.0
    m2 = Courtemanche(cfg)
    m2.define(state=m.state)                                                     # restart from the dict
    assert np.array_equal(m2._ctx.state['_Na_i_'], m.state['_Na_i_'])
    with pytest.raises(NotImplementedError):
        m3 = Courtemanche(cfg)
        m3.fast_states = ['V']
        m3.define()


def test_steps_per_launch_choice(recording, monkeypatch):
    """Fenton4v._steps_per_launch: two time steps per launch from 3072^2 cells up (with or without
    a phase field), never with a width that is not a multiple of 4; the config key and
    FIB_STEPS_PER_LAUNCH override the size rule (but not the restriction)."""
    monkeypatch.delenv('FIB_STEPS_PER_LAUNCH', raising=False)
    base = {'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5, 'duration': 1, 'timeline': False,
            'timeline_name': 'x', 'save_graph': False}

    def choice(width, height, hole=False, **extra):
        m = Fenton4v(dict(base, width=width, height=height, **extra))
        m._phase_rows = np.ones([2, 2], np.float32) if hole else None     # define() is not run: no arrays
        m._nranks = 1
        return m._steps_per_launch()

    assert choice(512, 512) == 1
    assert choice(4096, 4096) == 2
    assert choice(3072, 3072) == 2 and choice(3072, 3068) == 1
    assert choice(4096, 4096, hole=True) == 2 and choice(3072, 3072, hole=True) == 1
    assert choice(4098, 4096) == 1                                   # width % 4
    assert choice(4096, 4096, steps_per_launch=1) == 1
    assert choice(512, 512, steps_per_launch=2) == 2
    monkeypatch.setenv('FIB_STEPS_PER_LAUNCH', '1')
    assert choice(4096, 4096) == 1
    monkeypatch.setenv('FIB_STEPS_PER_LAUNCH', '2')
    assert choice(512, 512) == 2 and choice(512, 512, hole=True) == 2 and choice(514, 512) == 1
    assert choice(512, 512, steps_per_launch=1) == 1                 # the config key wins over the env
