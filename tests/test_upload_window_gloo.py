"""CPU, world_size 2 over gloo: the host-side agreement that lets NCCL shards step behind their pipelined
uploads (fib_tf_b200/ionic.py: _agree_on_upload_window, the _ctx property).  fib_step_behind_upload is a
collective with its own exchange pattern, so every rank has to issue it at the same point of its call sequence
with the same count -- or nobody does.  Exercised against the recording double of the C-ABI context: the double
computes nothing, it records what the host logic asks for."""
import os

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from test_run_loop_cpu import CFG, RecordingContext
from test_sharded_gloo import free_port


class ShardContext(RecordingContext):
    """+ the sharding / upload-session entry points; `upload` = what fib_upload_state reports on this rank"""
    upload = (False, False, 0, 0)

    def comm_init(self, nranks, rank, uid):
        self.calls.append(('comm', nranks, rank))

    def upload_state(self):
        return self.upload

    def step_behind_upload(self, n):
        self.calls.append(('behind', n))


def worker(rank, world, port, scenario, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from fib_tf_b200 import _capi
        from fib_tf_b200.fenton import Fenton4v
        _capi.Context = ShardContext
        _capi.comm_unique_id = lambda: bytes(128)
        good = (True, True, 1 if rank % 2 == 0 else -1, 7)
        ShardContext.upload = {
            'agree': good,
            'short_run': good,
            'one_rank_incomplete': good if rank == 0 else (True, False, -1, 0),
            'one_rank_wrong_direction': good if rank == 0 else (True, True, +1, 7),
            'early_touch': good,
            'no_upload': (False, False, 0, 0),
        }[scenario]
        iters = 3 if scenario == 'short_run' else 12
        m = Fenton4v(dict(CFG, distributed=True, duration=iters))       # 10 steps of 0.1 ms per iteration
        m.define()
        m.add_pace_op('s2', 'luq', 1.0)
        m._ctx.calls.clear()
        for i in m.run(None, block=False):
            if scenario == 'early_touch' and i == 2:
                m.fire_op('s2')                  # a collective call inside the window: the count so far goes first
        out[rank] = [c for c in m._ctx.calls if c[0] in ('behind', 'step', 'stim')]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('scenario', ['agree', 'short_run', 'one_rank_incomplete', 'one_rank_wrong_direction',
                                      'early_touch', 'no_upload'])
def test_ranks_agree_on_the_iterations_behind_the_upload(scenario):
    out = mp.Manager().dict()
    mp.spawn(worker, args=(2, free_port(), scenario, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    kinds = lambda calls: [(c[0], c[1] if c[0] == 'behind' else c[2] if c[0] == 'step' else None) for c in calls]
    assert kinds(a) == kinds(b), (a, b)          # the same sequence of collectives on both ranks
    seq = kinds(a)
    if scenario == 'agree':                      # min(room 7, window 10, 12 iterations): 7 behind, then 5 plain
        assert seq == [('behind', 7)] + [('step', 1)] * 5
    elif scenario == 'short_run':                # never more than the run has
        assert seq == [('behind', 3)]
    elif scenario in ('one_rank_incomplete', 'one_rank_wrong_direction', 'no_upload'):
        assert seq == [('step', 1)] * 12         # one rank cannot: nobody does
    else:                                        # iterations 0..2 run behind the upload BEFORE the stimulus lands
        assert seq == [('behind', 3), ('stim', None)] + [('step', 1)] * 9
