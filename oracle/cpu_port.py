"""TEST INFRASTRUCTURE (see oracle/README.md): ctypes loader for the plain-C + OpenMP port
(oracle/csrc/monodomain_cpu.c -> oracle/_build/libfiboracle.so) and for the compiled reference
header (oracle/_ref/libcourt_ref.so).  Used by tests/ and by bench.py's CPU-baseline legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, '_build', 'libfiboracle.so')
REF_SO = os.path.join(HERE, '_ref', 'libcourt_ref.so')
_FP = np.ctypeslib.ndpointer(dtype=np.float32, flags='C_CONTIGUOUS')


def build():
    subprocess.check_call(['make', '-s', '-C', HERE])


_port = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        L.fib_cpu_threads.restype = C.c_int
        L.fib_cpu_fenton_step.restype = None
        L.fib_cpu_fenton_step.argtypes = [C.c_int, C.c_int, _FP, _FP, _FP, _FP, _FP, C.c_void_p,
                                          C.c_double, C.c_double]
        L.fib_cpu_br_step.restype = None
        L.fib_cpu_br_step.argtypes = [C.c_int, C.c_int] + [_FP] * 9 + [C.c_void_p, C.c_double,
                                                                       C.c_double, C.c_int, C.c_int,
                                                                       C.c_void_p]
        _port = L
    return _port


class CPortModel:
    """OracleModel-compatible driver (define/iterate/state) on top of the C port; 4v and BR."""

    def __init__(self, kind, config):
        from . import monodomain_np as onp
        self._onp = onp
        self._np_model = onp.OracleModel(kind, config)      # reuses IC / phase / pace set-up only
        self.kind, self.cfg = kind, dict(config)
        if kind not in ('fenton4v', 'br'):
            raise NotImplementedError('C port covers fenton4v and br')

    def add_hole(self, *a):
        self._np_model.add_hole(*a)

    def define(self, s1=True, state=None):
        m = self._np_model
        m.define(s1)
        self.state = {k: np.ascontiguousarray(v) for k, v in m.state.items()}
        self.phase = None if m.phase is None else np.ascontiguousarray(m.phase, dtype=np.float32)
        self.cheb = (np.ascontiguousarray(m.coeffs, dtype=np.float32)
                     if m.coeffs is not None else None)
        self.dt_per_step = m.dt_per_step
        self._tmp = np.empty_like(self.state[m.POT[self.kind]])

    def add_pace(self, name, loc, v):
        self._np_model.add_pace(name, loc, v)

    def fire(self, name):
        if name == 'slow':
            return
        loc, v = self._np_model.paces[name]
        k = self._np_model.POT[self.kind]
        self.state[k] = np.ascontiguousarray(
            self._onp.apply_pace(self.state[k], loc, v, self._np_model.min_v))

    def pot(self):
        return self.state[self._np_model.POT[self.kind]]

    def iterate(self):
        L, s, c = port(), self.state, self.cfg
        H, W = self.pot().shape
        ph = None if self.phase is None else self.phase.ctypes.data_as(C.c_void_p)
        if self.kind == 'fenton4v':
            for _ in range(10):
                L.fib_cpu_fenton_step(H, W, s['U'], self._tmp, s['V'], s['W'], s['S'], ph,
                                      c['dt'], c['diff'])
                s['U'], self._tmp = self._tmp, s['U']
        else:
            sched = (5, 0, 0, 0, 0) if c.get('skip') else (1, 1, 1, 1, 1)
            cb = None if self.cheb is None else self.cheb.ctypes.data_as(C.c_void_p)
            for n in sched:
                L.fib_cpu_br_step(H, W, s['V'], self._tmp, s['C'], s['M'], s['H'], s['J'], s['D'],
                                  s['F'], s['XI'], ph, c['dt'], c['diff'], n,
                                  1 if c.get('cheby') else 0, cb)
                s['V'], self._tmp = self._tmp, s['V']


_ref = None


def court_ref():
    """The reference's own courtemanche.h, compiled (None where it was never built)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            if os.path.exists('/root/reference/courtemanche.h'):
                build()
            else:
                return None
        L = C.CDLL(REF_SO)
        L.ref_calc_inter.argtypes = [C.c_float, _FP]
        L.ref_init_table.argtypes = [_FP]
        L.ref_init_cell.argtypes = [_FP, C.c_int]
        L.ref_deriv.argtypes = [_FP, _FP, C.c_float, _FP, C.c_int]
        L.ref_euler_batch.argtypes = [_FP, _FP, C.c_long, C.c_int, C.c_float, _FP, C.c_int]
        for f in (L.ref_calc_inter, L.ref_init_table, L.ref_init_cell, L.ref_deriv, L.ref_euler_batch):
            f.restype = None
        _ref = L
    return _ref
