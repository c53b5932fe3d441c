"""GPU: BASELINE configs[0] / configs[1] end to end -- the reference's own 512^2, 1000-ms driver runs
(S1, hole, S2 -> a re-entrant spiral) -- against the UNMODIFIED reference executed under the TF shim
(oracle/make_golden_spiral.py -> tests/golden/spiral_*.npz).  After S2 the field is compared
STATISTICALLY, as BASELINE.json prescribes for long horizons: activation counts, cycle length
(rotation period) and action-potential duration at 8 probes within 1-2 %; before S2 the run is a
deterministic planar wave and activation times must agree to a fraction of an iteration."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


# Bounds, relative (round 2: 1 %, was 2-3 %).  profiles/r2_spiral_report.txt lists what is measured per probe --
# 4v / BR / court_ultra agree with the reference to <= 0.03 %, court.py to <= 0.52 % in APD -- next to how far the
# UNMODIFIED reference moves when only its math library is swapped (tests/golden/spiral_*_alt.npz,
# scripts/spiral_spread.py: <= 0.01 % for 4v, 0.9 % in APD / 0.45 ms in activation time for court.py).
CL_BOUND = 0.01
APD_BOUND = 0.01


def load(which):
    path = os.path.join(GOLDEN, 'spiral_%s.npz' % which)
    if not os.path.exists(path):
        pytest.skip('%s not generated' % path)
    z = np.load(path)
    return json.loads(str(z['meta'])), z['probes'], z['frames']


def events(trace, level, dt_iter):
    """(up-crossing times, down-crossing times) in ms with linear interpolation."""
    t = np.asarray(trace, np.float64)
    up, dn = [], []
    for i in range(1, len(t)):
        if t[i - 1] < level <= t[i]:
            up.append((i - 1 + (level - t[i - 1]) / (t[i] - t[i - 1])) * dt_iter)
        elif t[i - 1] >= level > t[i]:
            dn.append((i - 1 + (t[i - 1] - level) / (t[i - 1] - t[i])) * dt_iter)
    return np.array(up), np.array(dn)


def apds(up, dn):
    out = []
    for u in up:
        later = dn[dn > u]
        if len(later):
            out.append(later[0] - u)
    return np.array(out)


@pytest.mark.parametrize('which', ['fenton', 'br', 'court', 'court_ultra'])
def test_spiral_run_matches_the_reference_statistically(cuda_device, which):
    from fib_tf_b200.br import BeelerReuter
    from fib_tf_b200.court import Courtemanche
    from fib_tf_b200.court_ultra import Courtemanche as CourtemancheUltra
    from fib_tf_b200.fenton import Fenton4v
    meta, ref_probes, ref_frames = load(which)
    cls = {'fenton': Fenton4v, 'br': BeelerReuter, 'court': Courtemanche, 'court_ultra': CourtemancheUltra}
    model = cls[which](meta['config'])
    model.add_hole_to_phase_field(*meta['hole'])
    if meta.get('extra_hole'):
        model.add_hole_to_phase_field(*meta['extra_hole'])
    model.define()
    model.add_pace_op('s2', 'luq', meta['s2_value'])
    probes, frames = [], []
    name = model._pot_name
    for i in model.run(None):
        if which.startswith('court') and i % 10 == 0:
            model.fire_op('slow')                       # court.py:616-617
        if i == meta['s2_iter']:
            model.fire_op('s2')
        probes.append([model._ctx.probe(name, r, c) for r, c in meta['probes']])
        if i % meta['frame_every_iter'] == 0:
            frames.append(model._State[name].eval()[4::8, 4::8])
    probes, frames = np.asarray(probes, np.float32), np.asarray(frames)
    assert probes.shape == ref_probes.shape and np.isfinite(probes).all()
    assert model.nonfinite_cells() == {}
    model.close()

    lo, hi = {'fenton': (0.0, 1.0), 'br': (-90.0, 30.0)}.get(which, (-100.0, 50.0))
    level = lo + 0.5 * (hi - lo)
    dt_iter = meta['dt_per_step'] * meta['config']['dt']
    s2_ms = meta['s2_iter'] * dt_iter
    compared = 0
    report = []
    for k in range(probes.shape[1]):
        up_r, dn_r = events(ref_probes[:, k], level, dt_iter)
        up_c, dn_c = events(probes[:, k], level, dt_iter)
        # (a) deterministic phase: the S1 wave reaches every probe at the same time
        pre_r, pre_c = up_r[up_r < s2_ms], up_c[up_c < s2_ms]
        assert len(pre_r) == len(pre_c), (k, pre_r, pre_c)
        if len(pre_r):
            assert np.max(np.abs(pre_r - pre_c)) <= 0.25 * dt_iter + 2e-3 * pre_r.max(), (k, pre_r, pre_c)
        # (b) statistical phase
        assert abs(len(up_r) - len(up_c)) <= 1, (k, len(up_r), len(up_c))
        post_r, post_c = up_r[up_r > s2_ms + 100], up_c[up_c > s2_ms + 100]
        if len(post_r) >= 3 and len(post_c) >= 3:
            cl_r, cl_c = np.diff(post_r).mean(), np.diff(post_c).mean()
            a_r, a_c = apds(post_r, dn_r).mean(), apds(post_c, dn_c).mean()
            report.append('%s probe %d: cycle length %.3f vs %.3f ms (%.3f %%), APD %.3f vs %.3f ms (%.3f %%), %d beats'
                          % (which, k, cl_c, cl_r, 100 * abs(cl_c - cl_r) / cl_r, a_c, a_r, 100 * abs(a_c - a_r) / a_r,
                             len(post_r)))
            assert abs(cl_c - cl_r) <= CL_BOUND * cl_r, ('cycle length', k, cl_c, cl_r)
            assert abs(a_c - a_r) <= APD_BOUND * a_r, ('APD', k, a_c, a_r)
            compared += 1
        elif len(up_r) == len(up_c) and len(up_r) >= 2:
            # few beats in the window (Courtemanche, 700 ms): compare them one by one
            assert np.all(np.abs(up_c - up_r) <= 0.001 * up_r + 0.5), ('activation times', k, up_c, up_r)
            a_r, a_c = apds(up_r, dn_r), apds(up_c, dn_c)
            n = min(len(a_r), len(a_c))
            report.append('%s probe %d (beat by beat): activation times off by <= %.3f ms (%.3f %%), APD off by <= %.3f ms '
                          '(%.3f %%), %d beats' % (which, k, np.max(np.abs(up_c - up_r)),
                                                   100 * np.max(np.abs(up_c - up_r) / up_r),
                                                   np.max(np.abs(a_c[:n] - a_r[:n])) if n else 0.0,
                                                   100 * np.max(np.abs(a_c[:n] - a_r[:n]) / a_r[:n]) if n else 0.0, len(up_r)))
            assert len(a_r) == len(a_c) and np.all(np.abs(a_c - a_r) <= APD_BOUND * a_r), ('APD', k, a_c, a_r)
            compared += 1
    n_pre = int(meta['s2_iter'] // meta['frame_every_iter'])
    pre_err = float(np.max(np.abs(frames[:n_pre] - ref_frames[:n_pre]))) / (hi - lo)
    frac_err = max([abs(float((f_c > level).mean()) - float((f_r > level).mean()))
                    for f_c, f_r in zip(frames[n_pre + 2:], ref_frames[n_pre + 2:])] or [0.0])
    report.append('%s frames: before S2 point-wise within %.4f %% of the range, afterwards excited fraction within %.5f'
                  % (which, 100 * pre_err, frac_err))
    if os.environ.get('FIB_SPIRAL_REPORT'):
        with open(os.environ['FIB_SPIRAL_REPORT'], 'a') as f:
            f.write('\n'.join(report) + '\n')
    assert compared >= 3, 'too few probes could be compared'
    # frames before S2 agree point-wise within 1 % of the range (measured <= 0.28 %, court.py);
    # afterwards the excited fraction agrees
    assert pre_err <= 0.01, pre_err
    assert frac_err <= 0.01, frac_err


def test_lookup_table_flavour_tracks_the_exact_model(cuda_device):
    """The 150x30 table with its truncating 1-mV lookup (courtemanche.h:354-357) is an APPROXIMATION
    of calc_inter; this quantifies it on the court_ultra.py driver run: against the exact reference,
    every activation time within 3 % + 2 ms and every APD within 15 % (measured: <= 1.1 % on times,
    1-11 % on APD -- the truncation shortens the short, high-rate action potentials most)."""
    from fib_tf_b200.court_ultra import Courtemanche
    meta, ref_probes, _ = load('court_ultra')
    model = Courtemanche(dict(meta['config'], lut=True))
    model.add_hole_to_phase_field(*meta['hole'])
    model.add_hole_to_phase_field(*meta['extra_hole'])
    model.define()
    model.add_pace_op('s2', 'luq', meta['s2_value'])
    probes = []
    for i in model.run(None):
        if i == meta['s2_iter']:
            model.fire_op('s2')
        probes.append([model._ctx.probe('V', r, c) for r, c in meta['probes']])
    probes = np.asarray(probes, np.float32)
    model.close()
    dt_iter, compared = meta['config']['dt'], 0
    for k in range(probes.shape[1]):
        up_r, dn_r = events(ref_probes[:, k], -25.0, dt_iter)
        up_c, dn_c = events(probes[:, k], -25.0, dt_iter)
        if len(up_r) == len(up_c) and len(up_r) >= 2:
            assert np.all(np.abs(up_c - up_r) <= 0.03 * up_r + 2.0), (k, up_c, up_r)
            a_r, a_c = apds(up_r, dn_r), apds(up_c, dn_c)
            assert len(a_r) == len(a_c) and np.all(np.abs(a_c - a_r) <= 0.15 * a_r + 1.0), (k, a_c, a_r)
            compared += 1
    assert compared >= 4
