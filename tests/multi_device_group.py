"""Single process, several GPUs: fib_step_group over contexts living on DIFFERENT devices (halo rows
travel with cudaMemcpyPeerAsync) must be bit-identical to the unsharded run on device 0.
    python tests/multi_device_group.py        (needs >= 2 visible GPUs)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fib_tf_b200 import _capi  # noqa: E402
from fib_tf_b200.sharding import partition_rows  # noqa: E402


def main():
    ndev = _capi.device_count()
    if ndev < 2:
        print('needs >= 2 GPUs, found', ndev)
        return 0
    H, W, dt, diff = 1200, 1100, 0.1, 1.3
    rng = np.random.default_rng(0)
    init = {'U': rng.random((H, W), dtype=np.float32) * 0.3, 'V': np.ones((H, W), np.float32),
            'W': np.ones((H, W), np.float32), 'S': np.zeros((H, W), np.float32)}
    init['U'][:, :8] = 1.0
    whole = _capi.Context(_capi.FENTON4V, H, W, dt, diff, flags=_capi.F_NO_GRAPH, device=0)
    parts = partition_rows(H, ndev)
    shards = [_capi.Context(_capi.FENTON4V, H, W, dt, diff, flags=_capi.F_NO_GRAPH, device=d, row0=r0, rows=n)
              for d, (r0, n) in enumerate(parts)]
    for v, a in init.items():
        whole.set_state(v, a)
        for s, (r0, n) in zip(shards, parts):
            s.set_state(v, a[r0:r0 + n])
    whole.step(0, 3)
    _capi.step_group(shards, 0, 3)
    ok = True
    for v in init:
        got = np.concatenate([s.get_state(v) for s in shards], axis=0)
        same = np.array_equal(got, whole.get_state(v))
        ok &= same
        print('%s: %d devices, group == unsharded: %s' % (v, ndev, same))
    for c in shards + [whole]:
        c.close()
    print('MULTI-DEVICE GROUP OK' if ok else 'MULTI-DEVICE GROUP MISMATCH')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
