"""The reference's only published RESULT: conduction velocity vs `diff`
(diff_conduction_velcoty.dat:3-14).  One grid spacing is fitted per column (it is undocumented);
every row must then be reproduced within 1.5 % -- by the oracle on the CPU (this pins the oracle
against the reference's published numbers) and by the CUDA path on the GPU, which must also agree
with the oracle's CV within 1 % (BASELINE.json)."""
import numpy as np
import pytest

import cv_common as cv
from oracle import monodomain_np as onp


def oracle_cv(kind, diff, **flags):
    width = 420 if kind == 'fenton4v' else 300
    c1, c2 = (150, 350) if kind == 'fenton4v' else (100, 240)
    level = 0.5 if kind == 'fenton4v' else -30.0
    m = onp.OracleModel(kind, cv.strip_config(diff, width, **flags))
    m.define()
    dt_iter = m.dt_per_step * 0.1
    return cv.measure_cv(m, level, c1, c2, dt_iter, 4000, lambda mm: mm.pot()[2])


def test_oracle_reproduces_published_cv_4v():
    got = {d: oracle_cv('fenton4v', d) for d in (0.4, 0.8, 1.0, 1.5)}
    dx, worst = cv.fit_dx(got, cv.CV_TABLE_4V)
    assert abs(dx - 0.0302) < 0.0004, dx
    assert worst < 0.01, (dx, worst, got)


def test_oracle_reproduces_published_cv_br():
    got = {d: oracle_cv('br', d) for d in (0.4, 1.0, 2.0)}
    dx, worst = cv.fit_dx(got, cv.CV_TABLE_BR)
    assert 0.0290 < dx < 0.0310, dx
    assert worst < 0.02, (dx, worst, got)


@pytest.mark.gpu
@pytest.mark.parametrize('kind,table,flags', [
    ('fenton4v', cv.CV_TABLE_4V, {}),
    ('br', cv.CV_TABLE_BR, {}),
    ('br', cv.CV_TABLE_BR, {'cheby': True}),
])
def test_cuda_reproduces_published_cv_table(cuda_device, kind, table, flags):
    from cuda_adapter import CudaModel
    width = 700 if kind == 'fenton4v' else 520
    c1, c2 = (300, 600) if kind == 'fenton4v' else (200, 400)
    level = 0.5 if kind == 'fenton4v' else -30.0
    got = {}
    for diff in sorted(table):
        m = CudaModel(kind, cv.strip_config(diff, width, **flags))
        m.define()
        dt_iter = m.m.dt_per_step * 0.1
        got[diff] = cv.measure_cv(m, level, c1, c2, dt_iter, 6000,
                                  lambda mm: mm.m._ctx.get_rect(mm.m._pot_name, 2, 3, 0, width)[0])
        m.close()
    dx, worst = cv.fit_dx(got, table)
    # 4v: the north-star 1 % bar (measured worst residual 0.34 %).  The published BR column is itself
    # only self-consistent to ~1.5 % under ONE grid spacing (the CPU oracle
    # has the same 1.5 % residual) and cheby=True is another 0.6-2 % slower than the exact gates; the
    # 1 % CUDA-vs-reference bar for BR is test_cuda_cv_matches_oracle_within_1_percent below.
    tol = 0.01 if kind == 'fenton4v' else (0.02 if not flags else 0.035)
    assert 0.0290 < dx < 0.0312, dx
    assert worst < tol, (dx, worst, got)


@pytest.mark.gpu
@pytest.mark.parametrize('kind,diff,flags', [('fenton4v', 1.5, {}), ('br', 0.809, {'cheby': True}),
                                             ('br', 0.809, {'cheby': True, 'skip': True})])
def test_cuda_cv_matches_oracle_within_1_percent(cuda_device, kind, diff, flags):
    from cuda_adapter import CudaModel
    ref = oracle_cv(kind, diff, **flags)
    width = 420 if kind == 'fenton4v' else 300
    c1, c2 = (150, 350) if kind == 'fenton4v' else (100, 240)
    level = 0.5 if kind == 'fenton4v' else -30.0
    m = CudaModel(kind, cv.strip_config(diff, width, **flags))
    m.define()
    got = cv.measure_cv(m, level, c1, c2, m.m.dt_per_step * 0.1, 4000,
                        lambda mm: mm.m._ctx.get_rect(mm.m._pot_name, 2, 3, 0, width)[0])
    m.close()
    assert abs(got - ref) <= 0.01 * ref, (got, ref)
