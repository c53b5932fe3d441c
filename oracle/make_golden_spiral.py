#!/usr/bin/env python
"""
TEST INFRASTRUCTURE (see oracle/README.md).

make_golden_spiral.py -- BASELINE.json configs[0] / configs[1] end to end with the UNMODIFIED
reference under oracle/tfshim.py: the drivers of fenton.py:155-187 (4v 512^2, hole (256,256,30),
S2 'luq' at 210 ms, 1000 ms) and br.py:347-382 (BR 512^2, cheby=True, hole (150,200,40), S2 at
300 ms, 1000 ms).  Minutes of CPU time each, run once:

    python oracle/make_golden_spiral.py fenton      -> tests/golden/spiral_fenton.npz
    python oracle/make_golden_spiral.py br          -> tests/golden/spiral_br.npz
    python oracle/make_golden_spiral.py court | court_ultra   (700 ms, court.py / court_ultra.py loops)
    python oracle/make_golden_spiral.py <which> --alt   -> tests/golden/spiral_<which>_alt.npz: the same run
        with the facade's transcendental functions swapped for another correctly rounding library
        (tfshim.ALT_LIBM), probes only -- how far the REFERENCE ITSELF moves in cycle length / APD
        when nothing but its math library changes (the floor for any statistical bound)

Stored: the transmembrane variable at PROBES after every run() iteration ([samples, n_probes] fp32)
and 64x64 block-subsampled frames every 50 ms -- enough to compare rotation period, APD and
activation counts statistically (the dynamics are chaotic; planes cannot be compared point-wise
after the S2 stimulus).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
import oracle.tfshim as shim  # noqa: E402

shim.install()
import warnings  # noqa: E402

warnings.simplefilter('ignore')

PROBES = [(20, 256), (128, 128), (128, 384), (384, 128), (384, 384), (256, 400), (400, 256), (60, 60)]


def main(which, alt=False):
    shim.ALT_LIBM = bool(alt)
    if which == 'fenton':
        import fenton
        cfg = {'width': 512, 'height': 512, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5, 'duration': 1000,
               'timeline': False, 'timeline_name': 'unused.json', 'save_graph': False}
        model = fenton.Fenton4v(cfg)
        hole, s2_ms, s2_v = (256, 256, 30), 210, 1.0
    elif which == 'br':
        import br
        cfg = {'width': 512, 'height': 512, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 0.809, 'duration': 1000,
               'skip': False, 'cheby': True, 'timeline': False, 'timeline_name': 'unused.json',
               'save_graph': False}
        model = br.BeelerReuter(cfg)
        hole, s2_ms, s2_v = (150, 200, 40), 300, 10.0
    else:
        # court.py:582-621 / court_ultra.py:489-512 driver loops (holes, S2 'luq' = 10 mV, 'slow'
        # fired every 10th iteration), shortened to 700 ms
        import court
        import court_ultra
        ultra = which == 'court_ultra'
        cfg = {'width': 512, 'height': 512, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5 if ultra else 0.809,
               'duration': 700, 'skip': False, 'cheby': True, 'timeline': False,
               'timeline_name': 'unused.json', 'save_graph': False, 'ultra_slow': False}
        model = (court_ultra if ultra else court).Courtemanche(cfg)
        hole, s2_ms, s2_v = ((256, 256, 10) if ultra else (256, 256, 30)), (300 if ultra else 350), 10.0
        extra_hole = (256, 256, 250, True)
    model.add_hole_to_phase_field(*hole)
    if which.startswith('court'):
        model.add_hole_to_phase_field(*extra_hole)
    model.define()
    model.add_pace_op('s2', 'luq', s2_v)
    s2 = model.millisecond_to_step(s2_ms)
    every = model.millisecond_to_step(50)
    trace, frames = [], []
    t0 = time.time()
    out_stream = sys.stdout
    sys.stdout = open(os.devnull, 'w')
    try:
        for i in model.run(None):
            if which.startswith('court') and i % 10 == 0:
                model.fire_op('slow')
            if i == s2:
                model.fire_op('s2')
            x = model.pot().eval()
            trace.append([x[r, c] for r, c in PROBES])
            if i % every == 0:
                frames.append(x[4::8, 4::8].astype(np.float32))
            if i % 100 == 0:
                print('%s iteration %d  %.0f s' % (which, i, time.time() - t0), file=sys.stderr, flush=True)
    finally:
        sys.stdout.close()
        sys.stdout = out_stream
    meta = {'model': which, 'config': cfg, 'hole': hole, 'extra_hole': extra_hole if which.startswith('court') else None, 's2_ms': s2_ms, 's2_value': s2_v, 's2_iter': s2,
            'dt_per_step': model.dt_per_step, 'probes': PROBES, 'frame_every_iter': every,
            'generator': 'oracle/make_golden_spiral.py (unmodified reference under oracle/tfshim.py)',
            'seconds': time.time() - t0}
    if alt:
        meta['generator'] += ', tfshim.ALT_LIBM = True'
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'spiral_%s_alt.npz' % which),
                            meta=np.array(json.dumps(meta)), probes=np.asarray(trace, np.float32))
    else:
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'spiral_%s.npz' % which),
                            meta=np.array(json.dumps(meta)), probes=np.asarray(trace, np.float32),
                            frames=np.asarray(frames, np.float32))
    print('%s done in %.0f s' % (which, time.time() - t0))


if __name__ == '__main__':
    main(sys.argv[1], alt='--alt' in sys.argv[2:])
