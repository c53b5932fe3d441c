"""torchrun --nproc-per-node N tests/dist_parity.py : on N GPUs, the NCCL row-sharded run (one
process per GPU, halo rows over ncclSend/ncclRecv) must be BIT-IDENTICAL to the unsharded run
that rank 0 performs on its own GPU.  All three model families, with a phase field and a
stimulus that straddles the seams; a rank-local host write; two time steps per launch; and the
pipelined upload behind NCCL shards (fib_step_behind_upload through IonicModel.run())."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pipelined_upload_cases(CLASSES, base, local, rank, world):
    """Every rank uploads its shard block by block (even ranks top to bottom, odd ranks bottom to top) and run()
    steps its first iterations BEHIND the copies (fib_step_behind_upload: seam exchanges after every launch of
    the first / last block, none in between).  State and probe ring must equal upload-then-step, unsharded."""
    from fib_tf_b200 import _capi
    ok = True
    H, W, block = 160 * world, 200, 40
    for kind, extra, window, iters in (('fenton4v', {'steps_per_launch': 2}, 3, 6), ('fenton4v', {'steps_per_launch': 1}, 3, 5),
                                       ('br', {'cheby': True, 'skip': True}, 6, 8), ('court_ultra', {'ultra_slow': True}, 20, 30)):
        cfg = dict(base, width=W, height=H, **extra)
        m = CLASSES[kind](dict(cfg, distributed=True, device=local, upload_window=window))
        dt_iter = None
        models = [m]
        if rank == 0:
            models.append(CLASSES[kind](dict(cfg, device=local, steps_per_launch=1, persist=False)))
        for mm in models:
            mm.add_hole_to_phase_field(90, 60, 17)
            mm.define()
            mm.duration = iters * mm.dt_per_step * mm.dt + 1e-9
        rng = np.random.default_rng(1234)                      # the same planes on every rank
        names = m._ctx.var_names
        full = {}
        for v in names:
            base_v = m._State[v].eval()                        # (collective) the model's own initial state ...
            noise = rng.uniform(-1.0, 1.0, (H, W)).astype(np.float32)
            full[v] = (base_v + np.float32(0.02) * noise * np.maximum(np.abs(base_v), np.float32(1e-3))).astype(np.float32)
        row0, rows = m._row0, m._rows
        pinned = {v: _capi.pinned_empty((rows, W)) for v in names}
        for v in names:
            pinned[v][...] = full[v][row0:row0 + rows]
        starts = list(range(0, rows, block))
        for r in (starts if rank % 2 == 0 else reversed(starts)):       # block-major, every plane of a block
            for v in names:
                m._ctx_obj.set_rect_async(v, row0 + r, 0, pinned[v][r:min(r + block, rows)])
        if rank == 0:
            for v in names:
                models[1]._ctx.set_state(v, full[v])
        seen = []
        m.cl_observer = lambda i, cl: seen.append((i, cl))
        for i in m.run(None, block=False):
            pass
        used = m._upload_window_used
        if rank == 0:
            models[1]._ctx.step(0, iters)
        for v in names:
            got = m._State[v].eval()
            if rank == 0:
                same = np.array_equal(got, models[1]._State[v].eval(), equal_nan=True)
                ok &= same
                if not same:
                    print('MISMATCH pipelined %s %s' % (kind, v))
        if used != window:
            ok = False
            print('pipelined %s: window %d was not used (%d)' % (kind, window, used))
        if rank == 0:
            print('%-12s %s pipelined upload, %d iterations behind the copies, %d ranks: sharded == unsharded: %s'
                  % (kind, extra, used, world, ok), flush=True)
        for mm in models:
            mm.close()
        for v in names:
            _capi.pinned_free(pinned[v])
    return ok


def main():
    os.environ.setdefault('FIB_PIPELINE_MIN_CELLS', '0')
    os.environ.setdefault('FIB_PIPELINE_BLOCK_ROWS', '40')
    from cuda_adapter import CLASSES
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    base = {'width': 200, 'height': 131, 'dt': 0.1, 'dt_per_plot': 10, 'diff': 1.1, 'duration': 1,
            'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'skip': False,
            'cheby': False, 'ultra_slow': False}
    ok = True
    for kind, extra, iters in (('fenton4v', {}, 4), ('br', {'cheby': True, 'skip': True}, 5),
                               ('br', {'cheby': True, 'width': 1300, 'height': 900}, 3),
                               ('court', {}, 12), ('court_ultra', {'ultra_slow': True}, 8),
                               # two time steps per launch (two halo rows of all four planes per
                               # exchange) against ONE step per launch, unsharded; without and
                               # with a phase field
                               ('fenton4v', {'steps_per_launch': 2, 'hole': False}, 4),
                               ('fenton4v', {'steps_per_launch': 2, 'width': 1300, 'height': 900}, 3)):
        cfg = dict(base, **extra)
        hole = cfg.pop('hole', True)
        fused = cfg.get('steps_per_launch') == 2
        models = [CLASSES[kind](dict(cfg, distributed=True, device=local))]
        if rank == 0:
            models.append(CLASSES[kind](dict(cfg, device=local, steps_per_launch=1)))
        for m in models:
            if hole:
                m.add_hole_to_phase_field(90, 60, 17)
            m.define()
            m.add_pace_op('s2', 'luq', float(m.max_v) * 0.4)
        for i in range(iters):
            for m in models:
                m._ctx.step(0, 1)
                if kind == 'court' and i % 5 == 0:
                    m.fire_op('slow')
                if i == 1:
                    m.fire_op('s2')
            if i == 2:
                # a rank-LOCAL host write: a block on the LAST rows of rank 0's shard, written only by
                # rank 0 (fib_set_rect can only be called by the owner).  The neighbour's halo copy of
                # those rows is stale until the next fib_step refreshes it -- which must happen on every
                # rank without any rank-local knowledge (fib_capi.cu: refresh_halos_nccl).
                sh = models[0]
                pot = sh._pot_name
                blk = np.full((2, 40), float(sh.max_v) * 0.3, np.float32)
                r_edge = sh._rows - 2 if rank == 0 else None           # rank 0 owns rows [0, rows)
                if rank == 0:
                    sh._ctx.set_rect(pot, r_edge, 30, blk)
                    models[1]._ctx.set_rect(pot, r_edge, 30, blk)
        for name in models[0]._ctx.var_names:
            full = models[0]._State[name].eval()          # gathered over ranks (collective)
            if rank == 0:
                ref = models[1]._State[name].eval()
                same = np.array_equal(full, ref)
                ok &= same
                if not same:
                    print('MISMATCH %s %s max|d|=%g' % (kind, name, np.abs(full - ref).max()))
        if fused:
            # really two steps per launch: 5 x (top rows, bottom rows, interior) per iteration, not 10 x
            assert models[0]._ctx.launch_count() < 20 * iters, models[0]._ctx.launch_count()
        if rank == 0:
            print('%-12s %s %d ranks: sharded == unsharded: %s' % (kind, extra, world, ok), flush=True)
        for m in models:
            m.close()
    ok &= pipelined_upload_cases(CLASSES, base, local, rank, world)
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if not flag.item():
        sys.exit(1)
    if rank == 0:
        print('DIST PARITY OK')


if __name__ == '__main__':
    main()
