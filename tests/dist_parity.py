"""torchrun --nproc-per-node N tests/dist_parity.py : on N GPUs, the NCCL row-sharded run (one
process per GPU, halo rows over ncclSend/ncclRecv) must be BIT-IDENTICAL to the unsharded run
that rank 0 performs on its own GPU.  All three model families, with a phase field and a
stimulus that straddles the seams."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
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
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if not flag.item():
        sys.exit(1)
    if rank == 0:
        print('DIST PARITY OK')


if __name__ == '__main__':
    main()
