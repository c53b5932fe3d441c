"""CPU: host-side logic of the drop-in API (no GPU): sharding arithmetic, phase field, stimulus
regions, Chebyshev table, call-ordering errors, step arithmetic."""
import numpy as np
import pytest

from conftest import load_fixture
from fib_tf_b200.br import BeelerReuter
from fib_tf_b200.court import Courtemanche
from fib_tf_b200.fenton import Fenton4v
from fib_tf_b200.sharding import halo_plan, owner_of_row, partition_rows
from oracle import monodomain_np as onp

CFG = {'width': 56, 'height': 40, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.5, 'duration': 10,
       'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'skip': False, 'cheby': True,
       'ultra_slow': False}


@pytest.mark.parametrize('H,n', [(512, 1), (512, 8), (33, 4), (32768, 8), (7, 7), (10, 3)])
def test_partition_rows_covers_grid(H, n):
    parts = partition_rows(H, n)
    assert len(parts) == n and parts[0][0] == 0
    assert sum(r for _, r in parts) == H
    for (a0, ar), (b0, _br) in zip(parts, parts[1:]):
        assert a0 + ar == b0
    assert max(r for _, r in parts) - min(r for _, r in parts) <= 1
    for row in (0, H // 2, H - 1):
        r = owner_of_row(H, n, row)
        assert parts[r][0] <= row < parts[r][0] + parts[r][1]


def test_partition_rejects_too_many_ranks():
    with pytest.raises(ValueError):
        partition_rows(3, 4)


def test_halo_plan_seams_are_interior_borders_are_not():
    H, n = 40, 4
    plans = [halo_plan(H, n, r) for r in range(n)]
    assert len(plans[0]) == 1 and len(plans[-1]) == 1 and all(len(p) == 2 for p in plans[1:-1])
    # what r sends to r+1 is exactly the halo r+1 expects from r
    for r in range(n - 1):
        send = [p for p in plans[r] if p[0] == r + 1][0]
        recv = [p for p in plans[r + 1] if p[0] == r][0]
        assert send[1] == recv[2] and recv[1] == send[2]


def test_phase_field_matches_reference_fixture():
    meta, arr = load_fixture('court_multirate')
    m = Courtemanche(meta['config'])
    for h in meta['holes']:
        m.add_hole_to_phase_field(*h)
    assert m.phase.dtype == np.float32
    assert np.array_equal(m.phase, arr['phase'])


def test_hole_after_define_is_an_assertion_error():
    m = Fenton4v(CFG)
    m.defined = True
    with pytest.raises(AssertionError):
        m.add_hole_to_phase_field(1, 1, 1)


def test_pace_before_define_is_an_assertion_error():
    with pytest.raises(AssertionError):
        Fenton4v(CFG).add_pace_op('s2', 'luq', 1.0)


def test_pace_regions_match_reference_semantics():
    m = Fenton4v(CFG)
    m.defined = True
    H, W = CFG['height'], CFG['width']
    for loc in ('left', 'right', 'top', 'bottom', 'luq', 'llq', 'ruq', 'rlq'):
        m.add_pace_op(loc, loc, 1.0)
        r0, r1, c0, c1 = m._ops[loc][1]
        mine = np.zeros([H, W], np.float32)
        mine[r0:r1, c0:c1] = 1.0
        ref = onp.apply_pace(np.zeros([H, W], np.float32), loc, 1.0, 0.0)
        assert np.array_equal(mine, ref), loc


def test_config_keys_become_attributes_and_step_arithmetic():
    m = BeelerReuter(CFG)
    assert m.width == 56 and m.cheby is True and m.min_v == -90.0 and m.max_v == 30.0
    m.dt_per_step = 5
    assert m.millisecond_to_step(300) == 600
    f = Fenton4v(CFG)
    f.dt_per_step = 10
    assert f.millisecond_to_step(210) == 210
    with f as ctx:                      # dummy context manager (ionic.py:288-307)
        assert ctx is f and f.jit_scope() is f


def test_chebyshev_table_equals_oracle_and_reference_coefficients():
    m = BeelerReuter(CFG)
    t = m.chebyshev_table()
    assert t.shape == (12, 9) and np.array_equal(t, onp.br_cheby_coeffs())
    # T-basis coefficients of m_inf, and the integer basis rows
    a = BeelerReuter.monomial_basis(8)
    assert list(a[4]) == [1, 0, -4, 0, 1, 0, 0, 0, 0] and list(a[8][::2]) == [1, -16, 20, -8, 1]


def test_stencil_helpers_validate_before_touching_the_device():
    """IonicModel.laplace / enforce_boundary / phase_field / rush_larsen are eager device ops
    (fib_op_*); malformed arguments are refused on the host, before any CUDA call."""
    from fib_tf_b200 import _capi
    m = Fenton4v(CFG)
    with pytest.raises(_capi.FibError):
        m.laplace(None)
    with pytest.raises(_capi.FibError):
        m.enforce_boundary(np.zeros(5, np.float32))
    with pytest.raises(AssertionError):
        m.phase_field(np.zeros((CFG['height'] + 2, CFG['width'] + 2), np.float32))     # no phase field
    m.add_hole_to_phase_field(10, 10, 3)
    with pytest.raises(ValueError):
        m.laplace(np.zeros((7, 9), np.float32))          # plane shape != phase shape


def test_headless_screen_keeps_the_reference_protocol(tmp_path):
    from fib_tf_b200.screen import Screen
    im = Screen(4, 6, 'x', keep_every=2)
    for k in range(3):
        assert im.imshow(np.full([4, 6], 0.25 * k, np.float32))
    assert im.frames_shown == 3 and len(im.frames) == 2 and im.peek() and im.wait() is None
    with pytest.raises(ValueError):
        im.imshow(np.zeros([3, 3]))
    saved = im.save(str(tmp_path / 'frame.npy'))
    assert np.array_equal(np.load(saved), im.last)


def test_halo_plan_two_rows_deep_for_two_steps_per_launch():
    """depth = 2: each message is two consecutive rows; what a rank sends is what its neighbour
    expects, and the received rows are exactly the two rows beyond the shard."""
    H, n = 37, 4
    parts = partition_rows(H, n)
    plans = [halo_plan(H, n, r, depth=2) for r in range(n)]
    for r, plan in enumerate(plans):
        row0, rows = parts[r]
        for peer, send, recv in plan:
            assert row0 <= send and send + 2 <= row0 + rows              # I own what I send
            back = [p for p in plans[peer] if p[0] == r]
            assert len(back) == 1 and back[0][2] == send                  # the peer receives exactly that
            assert recv in (row0 - 2, row0 + rows)                        # just outside my block
    assert len(plans[0]) == 1 and len(plans[-1]) == 1 and all(len(p) == 2 for p in plans[1:-1])
    with pytest.raises(ValueError):
        halo_plan(5, 4, 3, depth=2)                                       # a 1-row shard cannot


def test_bench_upload_order_is_block_major_and_reversible():
    """bench.py uploads every plane of a 512-row strip before the next strip (so the library can step behind the
    copies), top to bottom -- or bottom to top on the odd ranks of a sharded run; either way each row of each plane
    exactly once, partial strips at the shard's edges included."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Rec:
        def __init__(self):
            self.calls = []

        def set_rect_async(self, name, r0, c0, block):
            self.calls.append((name, r0, block.shape[0]))

    W = 8
    strips = {v: [np.zeros((bench.TILE, W), np.float32), np.ones((bench.TILE, W), np.float32)] for v in 'UVWS'}
    row0, rows = 700, 1500                                  # starts and ends inside a strip
    up, down = Rec(), Rec()
    n_up = bench.upload_tiled(up, strips, row0, rows, W)
    n_down = bench.upload_tiled(down, strips, row0, rows, W, descending=True)
    assert n_up == n_down == 4 * rows * W * 4
    for rec, starts in ((up, [700, 1024, 1536, 2048]), (down, [2048, 1536, 1024, 700])):
        assert [c[1] for c in rec.calls[::4]] == starts
        for k in range(0, len(rec.calls), 4):               # the four planes of a strip together
            assert [c[0] for c in rec.calls[k:k + 4]] == list('UVWS') and len({c[1:] for c in rec.calls[k:k + 4]}) == 1
        covered = sorted((c[1], c[1] + c[2]) for c in rec.calls if c[0] == 'U')
        assert covered[0][0] == row0 and covered[-1][1] == row0 + rows
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
