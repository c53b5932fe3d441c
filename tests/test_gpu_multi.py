"""GPU, >= 2 devices (skipped otherwise): the multi-GPU paths inside pytest, so that the driver's own
GPU test run proves them.
  * tests/dist_parity.py under torchrun: one process per GPU, halo rows over NCCL send/recv, must be
    BIT-IDENTICAL to the unsharded run -- every model family, two time steps per launch, and a
    rank-LOCAL host write (only the owner rank calls fib_set_rect) followed by a step;
  * tests/multi_device_group.py: fib_step_group over contexts living on different devices of one
    process (halo rows by cudaMemcpyPeerAsync)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_nccl_row_shards_are_bit_identical_to_the_unsharded_run(cuda_device):
    n = _devices()
    if n < 2:
        pytest.skip('needs >= 2 GPUs, found %d' % n)
    n = 2 if n < 4 else 4
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'dist_parity.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = '\n'.join((r.stdout + r.stderr).splitlines()[-30:])
    assert r.returncode == 0 and 'DIST PARITY OK' in r.stdout, tail


def test_step_group_across_two_devices_of_one_process(cuda_device):
    if _devices() < 2:
        pytest.skip('needs >= 2 GPUs')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'multi_device_group.py')],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and 'MULTI-DEVICE GROUP OK' in r.stdout, (r.stdout + r.stderr)[-2000:]
