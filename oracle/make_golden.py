#!/usr/bin/env python
"""
TEST INFRASTRUCTURE (see oracle/README.md).

make_golden.py -- runs the UNMODIFIED reference (/root/reference/*.py) on the CPU
through oracle/tfshim.py and writes the golden fixtures tests/golden/*.npz.

Only runnable where /root/reference exists (the build container).  The fixtures
and this script are committed; the GPU box and the test-suite only read the
fixtures.  Re-run with:   python oracle/make_golden.py

Every case drives the reference exactly like its own __main__ blocks do
(fenton.py:155-187, br.py:347-382, court.py:582-636, court_ultra.py:489-527):

    model = Model(config); model.add_hole_to_phase_field(...); model.define()
    model.add_pace_op(name, loc, v)
    for i in model.run(None):
        if i % slow_every == 0: model.fire_op('slow'); model.fire_op('trend')   # Courtemanche
        if i == when:           model.fire_op(name)
        snapshot / probe

Fixture layout (np.savez_compressed):
    meta            json: model, config, holes, paces, slow_every, snaps, probe, vars
    phase           [H,W] fp32 phase field (absent when no hole)
    s{i}__{var}     [H,W] fp32 state plane `var` after the loop body of iteration i
    probe           [samples] fp32  pot()[probe_row, probe_col] after each iteration
    trend           [n,2] fp32   (court.py only)
    meta['noise']   {'s{i}__{var}': e}: rel_err (oracle/monodomain_np.py) between two runs of the
                    unmodified reference that differ ONLY in the fp32 math library used for
                    exp/expm1/log/tanh/pow (NumPy's vs correctly rounded) -- the libm noise.
    wide__{var}     [H,W] float64 planes of the last snapshot from the float64 run described next
    meta['rounding'] same metric between the fp32 reference and the SAME graph evaluated in float64
                    (fp32-rounded constants and inputs): the reference's total fp32 rounding error.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')

sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import oracle.tfshim as shim  # noqa: E402

shim.install()
import warnings  # noqa: E402

warnings.simplefilter('ignore')
import fenton  # noqa: E402
import br  # noqa: E402
import court  # noqa: E402
import court_ultra  # noqa: E402

MODELS = {
    'fenton4v': fenton.Fenton4v,
    'br': br.BeelerReuter,
    'court': court.Courtemanche,
    'court_ultra': court_ultra.Courtemanche,
}


def base_config(**kw):
    cfg = {
        'width': 56, 'height': 40, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5,
        'duration': 10, 'timeline': False, 'timeline_name': 'unused.json',
        'save_graph': False, 'skip': False, 'cheby': False, 'ultra_slow': False,
    }
    cfg.update(kw)
    return cfg


def state_of(model, kind):
    """name -> ndarray for every state variable of the running reference model."""
    if kind in ('court', 'court_ultra'):
        return {k: v.eval() for k, v in model._State.items()}
    want = ('U', 'V', 'W', 'S') if kind == 'fenton4v' else ('V', 'C', 'M', 'H', 'J', 'D', 'F', 'XI')
    out = {}
    for v in shim.VARIABLES:
        if v.name in want and v.name not in out:
            out[v.name] = v.eval()
    return out


def run_case(name, kind, cfg, holes=(), paces=(), slow_every=0, snaps=(), probe=None,
             s1=True, trend_op=False, _alt=None):
    from oracle.monodomain_np import rel_err, var_floor
    shim.ALT_LIBM = _alt == 'libm'
    shim.WIDE = _alt == 'wide'
    shim.reset_registry()
    model = MODELS[kind](cfg)
    for h in holes:
        model.add_hole_to_phase_field(*h)
    model.define(s1) if s1 is not True else model.define()
    for (_when, pname, loc, v) in paces:
        model.add_pace_op(pname, loc, v)
    out = {}
    trace = []
    trend = []
    t0 = time.time()
    sys.stdout = open(os.devnull, 'w')     # the reference prints 'elapsed', warnings
    try:
        for i in model.run(None):
            if slow_every and i % slow_every == 0:
                model.fire_op('slow')
                if trend_op:   # court.py:109-110 indexes [width//2, 20]: needs H > W//2
                    model.fire_op('trend')
                    trend.append(model._Trend.eval())
            for (when, pname, _loc, _v) in paces:
                if i == when:
                    model.fire_op(pname)
            if i in snaps:
                for k, a in state_of(model, kind).items():
                    out['s%d__%s' % (i, k)] = a if _alt == 'wide' else a.astype(np.float32)
            if probe is not None:
                trace.append(model.pot().eval()[probe[0], probe[1]])
    finally:
        sys.stdout.close()
        sys.stdout = sys.__stdout__
        shim.ALT_LIBM = False
        shim.WIDE = False
    if _alt:
        return out
    noise, rounding = {}, {}
    if probe is None:
        # second pass of the unmodified reference with the alternate fp32 libm: its deviation from
        # the first pass is the reference's own noise floor (see tfshim.ALT_LIBM)
        alt = run_case(name, kind, cfg, holes, paces, slow_every, snaps, probe, s1, trend_op,
                       _alt='libm')
        wide = run_case(name, kind, cfg, holes, paces, slow_every, snaps, probe, s1, trend_op,
                        _alt='wide')
        for k in out:
            if k.startswith('s') and '__' in k:
                fl = var_floor(kind, k.split('__', 1)[1])
                noise[k] = rel_err(alt[k], out[k], fl)
                rounding[k] = rel_err(wide[k], out[k], fl)
        # keep the float64 planes of the LAST snapshot: lets a test ask whether another fp32
        # implementation is as close to exact arithmetic as the reference's own fp32 run is
        last = 's%d__' % max(snaps)
        for k in list(out):
            if k.startswith(last):
                out['wide__' + k.split('__', 1)[1]] = np.asarray(wide[k], dtype=np.float64)
    meta = {
        'name': name, 'model': kind, 'config': cfg, 'holes': [list(h) for h in holes],
        'paces': [list(p) for p in paces], 'slow_every': slow_every,
        'snaps': sorted(snaps), 'probe': list(probe) if probe else None,
        'dt_per_step': model.dt_per_step, 'samples': model.samples, 's1': bool(s1),
        'vars': sorted(state_of(model, kind).keys()), 'noise': noise, 'rounding': rounding,
        'generator': 'oracle/make_golden.py (unmodified reference under oracle/tfshim.py)',
        'numpy': np.__version__,
    }
    out['meta'] = np.array(json.dumps(meta))
    if model.phase is not None:
        out['phase'] = np.asarray(model.phase, dtype=np.float32)
    if trace:
        out['probe'] = np.asarray(trace, dtype=np.float32)
    if trend:
        out['trend'] = np.asarray(trend, dtype=np.float32)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-22s %-12s %4dx%-4d iters=%-5d %6.1fs  %7.1f KB' % (
        name, kind, cfg['height'], cfg['width'], model.samples, time.time() - t0,
        os.path.getsize(path) / 1024.0))


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    # ---- short horizon, full planes (per-step parity: 100 time steps) ----
    if want('fenton_hole'):
        run_case('fenton_hole', 'fenton4v', base_config(duration=10, diff=1.5),
                 holes=[(28, 20, 6)], paces=[(5, 's2', 'luq', 1.0)], snaps={0, 4, 5, 9})
    if want('fenton_plain'):
        # odd sizes (W % 4 != 0), no phase field, every stimulus location incl. an unknown one
        run_case('fenton_plain', 'fenton4v', base_config(width=47, height=33, duration=10, diff=1.0),
                 paces=[(1, 'a', 'right', 1.0), (2, 'b', 'top', 0.7), (3, 'c', 'bottom', 1.0),
                        (4, 'd', 'llq', 0.9), (5, 'e', 'ruq', 1.0), (6, 'f', 'rlq', 0.5),
                        (7, 'g', 'left', 1.0), (8, 'h', 'nowhere', 1.0), (8, 'k', 'luq', 0.4)],
                 snaps={0, 1, 2, 3, 4, 5, 6, 7, 8, 9})
    for nm, ch, sk in (('br_exact', False, False), ('br_cheby', True, False),
                       ('br_skip', False, True), ('br_chebyskip', True, True)):
        if want(nm):
            run_case(nm, 'br', base_config(duration=10, diff=0.809, cheby=ch, skip=sk),
                     holes=[(15, 20, 6)], paces=[(10, 's2', 'luq', 10.0)], snaps={0, 9, 10, 19})
    if want('br_plain'):
        run_case('br_plain', 'br', base_config(width=45, height=31, duration=6, diff=1.3),
                 paces=[(3, 's2', 'rlq', 10.0)], snaps={0, 3, 11})
    cc = dict(width=40, height=24, diff=0.809)
    if want('court_multirate'):
        run_case('court_multirate', 'court', base_config(duration=10, **cc),
                 holes=[(20, 12, 4), (20, 12, 18, True)], paces=[(50, 's2', 'luq', 10.0)],
                 slow_every=10, snaps={0, 9, 10, 49, 50, 99}, trend_op=True)
    if want('court_ultra'):
        run_case('court_ultra', 'court_ultra', base_config(duration=10, **dict(cc, diff=1.5)),
                 holes=[(20, 12, 4)], paces=[(50, 's2', 'luq', 10.0)],
                 slow_every=10, snaps={0, 9, 49, 50, 99})
    if want('court_ultra_us'):
        run_case('court_ultra_us', 'court_ultra',
                 base_config(duration=10, ultra_slow=True, **dict(cc, diff=1.5)),
                 paces=[(50, 's2', 'llq', 10.0)], slow_every=10, snaps={0, 9, 50, 99})

    # ---- long horizon on a thin strip (one full action potential; APD / CV) ----
    strip = dict(width=96, height=5)
    if want('fenton_long'):
        run_case('fenton_long', 'fenton4v', base_config(duration=450, diff=1.5, **strip),
                 snaps={99, 449}, probe=(2, 48))
    if want('br_long_exact'):
        run_case('br_long_exact', 'br', base_config(duration=450, diff=0.809, **strip),
                 snaps={199, 899}, probe=(2, 48))
    if want('br_long_cheby'):
        run_case('br_long_cheby', 'br', base_config(duration=450, diff=0.809, cheby=True, **strip),
                 snaps={199, 899}, probe=(2, 48))
    if want('br_long_chebyskip'):
        run_case('br_long_chebyskip', 'br',
                 base_config(duration=450, diff=0.809, cheby=True, skip=True, **strip),
                 snaps={199, 899}, probe=(2, 48))
    if want('court_long'):
        run_case('court_long', 'court', base_config(duration=400, diff=0.809, width=64, height=5),
                 slow_every=10, snaps={999, 3999}, probe=(2, 40))
    if want('court_ultra_long'):
        run_case('court_ultra_long', 'court_ultra',
                 base_config(duration=400, diff=1.5, width=64, height=5),
                 slow_every=10, snaps={999, 3999}, probe=(2, 40))


if __name__ == '__main__':
    main()
