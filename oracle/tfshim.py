"""
TEST INFRASTRUCTURE (see oracle/README.md) -- never imported by fib_tf_b200/.

tfshim: a lazy-graph NumPy stand-in for the TensorFlow-1.x API subset used by
the reference (/root/reference/{ionic,fenton,br,court,court_ultra}.py), so the
reference's own define()/run()/solve()/fire_op() run UNMODIFIED on the CPU and
produce golden vectors (oracle/make_golden.py).

Semantics reproduced (what TF 1.x does, restated; nothing copied):
  * graph mode: every tf.* call builds a Node; nothing is evaluated until
    Session.run(op) / .eval().  Session.run of a group of assigns evaluates all
    new values from the *current* variable values first and assigns afterwards
    (the dependency structure the reference relies on).
  * dtype: a Python scalar / np.float64 operand next to a float32 tensor is
    converted to float32 first (TF: convert_to_tensor(x, dtype=tensor.dtype)),
    so all arithmetic stays fp32 -- this plugs the two NumPy-2 fp64 leaks named
    in SURVEY.md section 8(c): br.py:327-331 (np.float64 Chebyshev coefficients)
    and court.py:189,243 (Python-scalar tau into tf.expm1).
  * tf.pad modes REFLECT / SYMMETRIC == numpy 'reflect' / 'symmetric'.
  * tf.contrib is absent, so IonicModel.jit_scope() (ionic.py:294-307) falls
    back to the model itself as a dummy context, exactly as it does without XLA.

Usage:  import oracle.tfshim as shim; shim.install(); import fenton  (with
/root/reference on sys.path).  shim.VARIABLES lists every tf.Variable created,
in creation order, with its name.
"""
import contextlib
import sys
import types

import numpy as np

F32 = np.float32
VARIABLES = []          # every Variable created since the last reset_registry()

# WIDE: run the whole graph in float64 -- constants and inputs are still rounded to float32 first
# (the model keeps the reference's fp32 parameters), only the ARITHMETIC is carried out in double.
# |reference(fp32) - reference(WIDE)| is the reference's own total fp32 rounding error: the level
# below which agreement between two different fp32 implementations of it cannot be demanded.
WIDE = False


def reset_registry():
    del VARIABLES[:]


def _coerce(x):
    """What TF makes of a non-tensor operand that meets a float32 tensor."""
    if isinstance(x, (bool, np.bool_)):
        return x
    if isinstance(x, (int, float, np.integer, np.floating)):
        return np.float64(F32(x)) if WIDE else F32(x)
    a = np.asarray(x)
    if a.dtype == np.float64 and not WIDE:
        return a.astype(F32)
    if WIDE and a.dtype == np.float32:
        return a.astype(np.float64)
    return a


class Node:
    """One op of the lazy graph."""
    __array_ufunc__ = None      # np.float64 * Node must defer to Node.__rmul__

    def __init__(self, fn, args):
        self.fn = fn
        self.args = args

    def value(self, memo):
        k = id(self)
        if k not in memo:
            vals = [a.value(memo) if isinstance(a, Node) else a for a in self.args]
            memo[k] = self.fn(*vals)
        return memo[k]

    def eval(self):
        return np.array(self.value({}))

    # arithmetic ----------------------------------------------------------
    def _bin(self, other, uf, swap=False):
        a, b = (other, self) if swap else (self, other)
        return Node(lambda x, y: uf(_coerce(x), _coerce(y)), [a, b])

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    def __lt__(self, o): return self._bin(o, np.less)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    def __neg__(self): return Node(lambda x: np.negative(_coerce(x)), [self])
    __hash__ = object.__hash__

    def __getitem__(self, idx):
        return _Slice(self, idx)

    def assign(self, value):
        return assign(self, value)


class _Slice(Node):
    def __init__(self, base, idx):
        Node.__init__(self, lambda a: a[idx], [base])
        self.base = base
        self.idx = idx


class Variable(Node):
    def __init__(self, initial_value, name=None, dtype=None):
        Node.__init__(self, None, [])
        self.init = np.array(initial_value)
        if WIDE and self.init.dtype == np.float32:
            self.init = self.init.astype(np.float64)
        self.val = self.init.copy()
        self.name = name
        VARIABLES.append(self)

    def value(self, memo):
        return self.val

    def eval(self):
        return self.val.copy()


class _Assign:
    def __init__(self, ref, value):
        self.ref = ref
        self.value = value

    def run(self):
        _run(self)


class _Group:
    def __init__(self, ops):
        self.ops = list(ops)

    def run(self):
        _run(self)


def _flatten(op, out):
    if isinstance(op, _Group):
        for o in op.ops:
            _flatten(o, out)
    elif isinstance(op, _Assign):
        out.append(op)
    elif op is not None:
        raise TypeError('tfshim: cannot run %r' % (op,))
    return out


def _run(op):
    if isinstance(op, Node):
        return op.eval()
    memo = {}
    todo = _flatten(op, [])
    # evaluate everything from the current variable values, then assign
    new = [(a.ref, np.array(a.value.value(memo) if isinstance(a.value, Node)
                            else _coerce(a.value))) for a in todo]
    for ref, v in new:
        if isinstance(ref, Variable):
            ref.val = np.broadcast_to(v, ref.val.shape).astype(ref.val.dtype)
        elif isinstance(ref, _Slice) and isinstance(ref.base, Variable):
            nv = ref.base.val.copy()
            nv[ref.idx] = v
            ref.base.val = nv
        else:
            raise TypeError('tfshim: assign target must be a Variable (slice)')
    return None


# ---- the tf.* surface --------------------------------------------------
def assign(ref, value, name=None):
    return _Assign(ref, value)


def group(*ops, **kw):
    return _Group(ops)


def constant(v, dtype=None, name=None):
    a = np.asarray(v)
    if a.dtype == np.float64:
        a = a.astype(F32)
    return Node(lambda: _coerce(a), [])


# ALT_LIBM: evaluate the transcendental functions in float64 and round once to float32, i.e. run
# the reference with a DIFFERENT but equally legitimate (in fact correctly rounded) fp32 math
# library.  |reference(ALT_LIBM) - reference(NumPy fp32 libm)| is the reference's own sensitivity
# to <= 1-ulp differences in exp/expm1/log/tanh/pow -- the noise floor below which "parity" with
# any other fp32 implementation (TF/Eigen, XLA, CUDA) has no meaning (see oracle/make_golden.py).
ALT_LIBM = False
_TRANSCENDENTAL = {np.exp, np.expm1, np.log, np.tanh}


def _apply(uf, a):
    a = _coerce(a)
    if ALT_LIBM and uf in _TRANSCENDENTAL:
        with np.errstate(all='ignore'):
            return uf(np.asarray(a, dtype=np.float64)).astype(F32)
    return uf(a)


def _unary(uf):
    def f(x, name=None):
        if isinstance(x, Node):
            return Node(lambda a: _apply(uf, a), [x])
        return Node(lambda: _apply(uf, x), [])
    return f


exp = _unary(np.exp)
expm1 = _unary(np.expm1)
log = _unary(np.log)
tanh = _unary(np.tanh)
sign = _unary(np.sign)
sqrt = _unary(np.sqrt)
square = _unary(np.square)
reciprocal = _unary(np.reciprocal)
abs = _unary(np.abs)            # noqa: A001  (mirrors tf.abs)


def _pow(a, b):
    a, b = _coerce(a), _coerce(b)
    if ALT_LIBM:
        return np.power(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)).astype(F32)
    return np.power(a, b)


def pow(x, y, name=None):       # noqa: A001
    return Node(_pow, [x, y])


# min / max follow the reference's GPU target (Eigen's CUDA mini/maxi are fminf/fmaxf; XLA:GPU used
# minnum/maxnum): the non-NaN operand wins.  np.fmin / np.fmax have exactly those semantics.
def maximum(x, y, name=None):
    return Node(lambda a, b: np.fmax(_coerce(a), _coerce(b)), [x, y])


def minimum(x, y, name=None):
    return Node(lambda a, b: np.fmin(_coerce(a), _coerce(b)), [x, y])


def where(cond, x, y, name=None):
    return Node(lambda c, a, b: np.where(c, _coerce(a), _coerce(b)), [cond, x, y])


def clip_by_value(x, lo, hi, name=None):
    # TF 1.x: maximum(minimum(t, clip_value_max), clip_value_min)
    return Node(lambda a: np.fmax(np.fmin(_coerce(a), _coerce(hi)), _coerce(lo)), [x])


def pad(x, paddings, mode='CONSTANT', name=None):
    def f(a, p):
        p = [tuple(int(q) for q in row) for row in np.asarray(p)]
        return np.pad(a, p, mode=mode.lower())
    return Node(f, [x, paddings])


@contextlib.contextmanager
def device(name):
    yield


@contextlib.contextmanager
def name_scope(name):
    yield name


class Session:
    graph = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, op, options=None, run_metadata=None):
        return _run(op)


class _Init:
    def run(self):
        for v in VARIABLES:
            v.val = v.init.copy()


def global_variables_initializer():
    return _Init()


def install():
    """Put the shim (and stubs for screen / timeline) into sys.modules."""
    me = sys.modules[__name__]
    tf = types.ModuleType('tensorflow')
    for k in ('assign group constant exp expm1 log tanh sign sqrt square reciprocal abs '
              'pow maximum minimum where clip_by_value pad device name_scope Session '
              'global_variables_initializer Variable').split():
        setattr(tf, k, getattr(me, k))
    tf.float32 = F32
    tf.__shim__ = True
    py = types.ModuleType('tensorflow.python')
    cl = types.ModuleType('tensorflow.python.client')
    tl = types.ModuleType('tensorflow.python.client.timeline')
    cl.timeline = tl
    py.client = cl
    tf.python = py
    sys.modules['tensorflow'] = tf
    sys.modules['tensorflow.python'] = py
    sys.modules['tensorflow.python.client'] = cl
    sys.modules['tensorflow.python.client.timeline'] = tl
    scr = types.ModuleType('screen')        # court.py:27 imports it at module level
    scr.Screen = type('Screen', (), {})
    sys.modules.setdefault('screen', scr)
    if not hasattr(np, 'int'):              # br.py:319 uses the removed alias
        np.int = int
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 200000))
    return tf
