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


def halo_plan(height, nranks, rank):
    """What rank exchanges after every time step: list of (peer, send_row, recv_halo) in GLOBAL
    rows; global rows 0 and H-1 are physical borders, shard seams are ordinary interior."""
    row0, rows = partition_rows(height, nranks)[rank]
    plan = []
    if rank > 0:
        plan.append((rank - 1, row0, row0 - 1))
    if rank + 1 < nranks:
        plan.append((rank + 1, row0 + rows - 1, row0 + rows))
    return plan
