"""CPU, world_size 2 over gloo: the N>1 host path.  Each rank owns a row block of the grid; the
NumPy oracle is stepped on the block plus TWO halo rows per seam (the outer one absorbs the
oracle's border treatment), halo rows travel with torch.distributed send/recv following
fib_tf_b200.sharding.halo_plan, and the result must be BIT-IDENTICAL to the unsharded oracle.
This pins the decomposition semantics the CUDA path implements (seams are interior; only global
rows 0 and H-1 are borders) and the rank plumbing (unique-id style broadcast)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fib_tf_b200.sharding import halo_plan, partition_rows
from oracle import monodomain_np as onp

H, W, STEPS = 37, 29, 12


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def exchange(local, rank, world, halo=2):
    """local: dict var -> [rows(+halos), W]; refresh `halo` rows per seam from the neighbours, following
    the product's exchange plan (fib_tf_b200.sharding.halo_plan: peer, first row sent, first halo row
    received, in GLOBAL rows -- the same plan the NCCL path of libfibb200 implements)."""
    row0, rows = partition_rows(H, world)[rank]
    lo = max(row0 - halo, 0)                       # global row of local row 0
    plan = halo_plan(H, world, rank, depth=halo)
    for name in sorted(local):
        a = local[name]
        reqs, recvs = [], []
        for peer, send_row, recv_row in plan:
            send = np.ascontiguousarray(a[send_row - lo:send_row - lo + halo])
            buf = torch.empty(halo, W)
            reqs.append(dist.isend(torch.from_numpy(send), peer))
            reqs.append(dist.irecv(buf, peer))
            recvs.append((recv_row - lo, buf))
        for r in reqs:
            r.wait()
        for at, buf in recvs:
            a[at:at + halo] = buf.numpy()


def worker(rank, world, port, out, every=1):
    """every = time steps between halo exchanges: 1 = the one-step kernels (one row of the diffusing
    variable would do; all planes are sent here for simplicity), 2 = two time steps per launch
    (csrc/fib_fused.cuh): ALL planes, `every` rows deep, the first step recomputed in the halo.
    The oracle treats the outermost local row as a border ring, which costs one more halo row than
    the CUDA kernels need: halo = every + 1."""
    halo = every + 1
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    box = [b'unique-id-from-rank-0' if rank == 0 else None]     # the ncclUniqueId plumbing
    dist.broadcast_object_list(box, src=0)
    assert box[0] == b'unique-id-from-rank-0'
    row0, rows = partition_rows(H, world)[rank]
    full = onp.fenton_init(H, W)
    phase = onp.hole_phase(None, H, W, 12, 15, 5)
    lo, hi = max(row0 - halo, 0), min(row0 + rows + halo, H)
    local = {k: v[lo:hi].copy() for k, v in full.items()}
    ph = phase[lo:hi]
    top = row0 - lo
    for step in range(STEPS):
        new = onp.fenton_step(local, 0.1, 1.5, ph)
        local = {k: np.array(v) for k, v in new.items()}
        if (step + 1) % every == 0:
            exchange(local, rank, world, halo)
    np.save(out % rank, np.stack([local[k][top:top + rows] for k in ('U', 'V', 'W', 'S')]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('every', [1, 2])
def test_two_rank_row_sharding_matches_unsharded_oracle(tmp_path, every):
    world = 2
    out = str(tmp_path / 'rank%d.npy')
    mp.spawn(worker, args=(world, free_port(), out, every), nprocs=world, join=True)
    st = onp.fenton_init(H, W)
    phase = onp.hole_phase(None, H, W, 12, 15, 5)
    for _ in range(STEPS):
        st = onp.fenton_step(st, 0.1, 1.5, phase)
    got = np.concatenate([np.load(out % r) for r in range(world)], axis=1)
    for i, k in enumerate(('U', 'V', 'W', 'S')):
        assert np.array_equal(got[i], st[k]), k
