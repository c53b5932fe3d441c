"""
fib_tf_b200.fenton -- drop-in for the reference's fenton.py: the Cherry-Ehrlich-Nattel-Fenton
(4v) canine left-atrial model (Heart Rhythm 2007;4(12):1553-62).

Same class name, constructor, define(s1), pot(), image() and driver block as fenton.py:31-187.
The ten unrolled explicit-Euler steps of one run() iteration (fenton.py:133-138) are ten
launches of the fused sm_100a kernel (boundary + 9-point Laplacian + phase term + 4v reaction),
replayed as one CUDA graph.
"""
import numpy as np

from . import _capi
from .ionic import DeviceVar, IonicModel


class Fenton4v(IonicModel):
    MODEL_ID = _capi.FENTON4V
    _pot_name = 'U'

    def __init__(self, props):
        IonicModel.__init__(self, props)
        self.min_v = 0.0
        self.max_v = 1.0
        self.depol = 0.0

    def define(self, s1=True):
        """Initial state U=0, V=W=1, S=0; S1 = column 1 of U set to 1 (fenton.py:116-123)."""
        super().define()
        ctx = self._make_context()
        init = {'U': 0.0, 'V': 1.0, 'W': 1.0, 'S': 0.0}
        for name, val in init.items():
            a = self._local_full(val)
            if s1 and name == 'U':
                a[:, 1] = 1.0
            ctx.set_state(name, a)
        self.dt_per_step = ctx.dt_per_step      # 10
        self._ode_op = _capi.OP_ODE
        self._State = {n: DeviceVar(self, n) for n in ctx.var_names}
        self._U = self._State['U']

    def pot(self):
        return self._U

    def image(self):
        return self._U.eval()


if __name__ == '__main__':
    # the reference's driver (fenton.py:155-187): 512^2, hole, S1-S2 at 210 ms, a frame every 10 ms
    model = Fenton4v({'width': 512, 'height': 512, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5,
                      'duration': 1000, 'timeline': False, 'timeline_name': 'timeline_4v.json',
                      'save_graph': False})
    model.add_hole_to_phase_field(256, 256, 30)
    model.define()
    model.add_pace_op('s2', 'luq', 1.0)
    s2, every = model.millisecond_to_step(210), model.millisecond_to_step(10)
    cube = np.zeros([int(model.duration / 10.0), model.height, model.width], dtype=np.float32)
    for i in model.run(None):           # headless; pass a fib_tf_b200.screen.Screen to watch
        if i == s2:
            model.fire_op('s2')
        if i % every == 0:
            cube[i // every] = model.image() * model.phase
    np.save('cube', cube)
