"""Row partition of the grid over ranks: contiguous row blocks, so the halo
rows of the row-major planes are contiguous (4*W bytes each).  Pure host logic."""


def partition_rows(height, nranks):
    """[(row0, rows)] for rank 0..nranks-1: contiguous, covering [0, height), sizes differ by at
    most one row (the first `height % nranks` ranks get the extra row)."""
    if nranks < 1:
        raise ValueError('nranks must be >= 1')
    if height < nranks:
        raise ValueError('cannot shard %d rows over %d ranks' % (height, nranks))
    base, extra = divmod(height, nranks)
    out, row = [], 0
    for r in range(nranks):
        n = base + (1 if r < extra else 0)
        out.append((row, n))
        row += n
    return out


def owner_of_row(height, nranks, row):
    """Rank owning global row `row`."""
    for r, (r0, n) in enumerate(partition_rows(height, nranks)):
        if r0 <= row < r0 + n:
            return r
    raise IndexError(row)


def halo_plan(height, nranks, rank, depth=1):
    """What rank exchanges after every launch: list of (peer, send_row, recv_halo) in GLOBAL rows;
    global rows 0 and H-1 are physical borders, shard seams are ordinary interior.

    depth = time steps per launch = rows per message: 1 -> one row of the diffusing variable per
    step; 2 -> with two time steps per launch (csrc/fib_fused.cuh) `send_row` and `recv_halo` are
    the FIRST of `depth` consecutive rows, of every state plane (the first step is recomputed on
    the neighbour's edge rows, which needs all of its variables)."""
    row0, rows = partition_rows(height, nranks)[rank]
    if depth < 1 or (nranks > 1 and rows < depth):
        raise ValueError('a shard of %d rows cannot exchange %d-row halos' % (rows, depth))
    plan = []
    if rank > 0:
        plan.append((rank - 1, row0, row0 - depth))
    if rank + 1 < nranks:
        plan.append((rank + 1, row0 + rows - depth, row0 + rows))
    return plan
