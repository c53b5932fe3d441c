"""
fib_tf_b200.egm -- drop-in for the reference's egm.py: a Beeler-Reuter run whose two Gaussian-mask
pseudo-electrograms are sampled every millisecond and written to test.dat.  The reference pulls the
full frame to the host for every sample (np.mean(model.image() * mask), egm.py:44-47); here each
sample is one weighted reduction on the device (fib_masked_sum).
"""
import numpy as np

from .br import BeelerReuter


def create_mask(model, x, y, radius):
    """Circular Gaussian mask centred at (x, y) (egm.py:5-12)."""
    xx, yy = np.meshgrid(np.arange(model.width), np.arange(model.height))
    dist = np.hypot(xx - x, yy - y)
    return np.array(np.exp(-(dist / radius) ** 2), dtype=np.float32)


def run(config, out='test.dat'):
    model = BeelerReuter(config)
    model.add_hole_to_phase_field(150, 256, 50)
    model.define()
    model.add_pace_op('s2', 'luq', 10.0)
    s2 = model.millisecond_to_step(300)
    m1 = model.add_probe_mask(create_mask(model, 300 + 15, 256, 5))
    m2 = model.add_probe_mask(create_mask(model, 300 - 15, 256, 5))
    trace = []
    every = max(int(10 / model.dt_per_step), 1)        # every 1 ms
    for i in model.run(None):
        if i == s2:
            model.fire_op('s2')
        if i % every == 0:
            trace.append([model.masked_image_mean(m1), model.masked_image_mean(m2)])
    trace = np.asarray(trace)
    if out:
        np.savetxt(out, trace)
    return trace


if __name__ == '__main__':
    run({'width': 512, 'height': 512, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.0, 'duration': 3000,
         'skip': False, 'cheby': True, 'timeline': False, 'timeline_name': 'timeline_br.json',
         'save_graph': False})
