#!/usr/bin/env python
"""
TEST INFRASTRUCTURE (see oracle/README.md).

facade_baseline.py -- BASELINE.md section 3.2: time the UNMODIFIED reference's own CPU execution of
BASELINE config 1 (fenton.py: Fenton 4v 512x512, hole (256,256,30), dt 0.1, diff 1.5) through the NumPy
TensorFlow facade (oracle/tfshim.py), one core, >= 200 time steps.  /root/reference exists only in the
build container, so this runs HERE and the result is committed as profiles/r2_facade_baseline.json;
bench.py quotes it as cpu_baseline.numpy_restatement.facade_recorded next to the same measurement of
the oracle's NumPy restatement made on the GPU box.

    python oracle/facade_baseline.py [iterations]
"""
import json
import os
import platform
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
import oracle.tfshim as shim  # noqa: E402

shim.install()
import warnings  # noqa: E402

warnings.simplefilter('ignore')


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    import fenton
    cfg = {'width': 512, 'height': 512, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5,
           'duration': (iters + 2) * 1.0, 'timeline': False, 'timeline_name': 'unused.json', 'save_graph': False}
    model = fenton.Fenton4v(cfg)
    model.add_hole_to_phase_field(256, 256, 30)
    model.define()
    t0 = None
    done = 0
    for i in model.run(None):               # the reference's own run() generator: one Session.run per iteration
        if i == 1:                          # two warm-up iterations
            t0 = time.perf_counter()
        elif i > 1:
            done += 1
    sec = time.perf_counter() - t0
    steps = done * model.dt_per_step
    out = {'value': 512 * 512 * steps / sec / 1e9, 'unit': 'Gcell-steps/s', 'cores': 1, 'kind': 'reference',
           'what': "the unmodified reference fenton.py (define() + run()) executed by oracle/tfshim.py's NumPy "
                   'TensorFlow facade: BASELINE config 1, %d time steps in %.1f s' % (steps, sec),
           'where': 'build container: %s, %d logical CPUs' % (platform.processor() or platform.machine(), os.cpu_count()),
           'published_by_reference': '0.052 Gcell-steps/s: 50 s per 10 000 steps of 512^2 on a 1.7 GHz quad-core, '
                                     'TensorFlow CPU device (details.md:264)'}
    path = os.path.join(ROOT, 'profiles', 'r2_facade_baseline.json')
    json.dump(out, open(path, 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
