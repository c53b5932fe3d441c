"""CPU: the JSON-line contract of `bench.py --impl reference` (the arm the driver runs on the GPU
box's host cores): every key of the base contract, the reference-arm additions, one line, and --
under a multi-rank launch -- silence from every rank but 0.  The timed thing here is the oracle's
C/OpenMP port (test infrastructure), which is exactly what that arm is allowed to execute."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEED = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
        'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline', 'impl']


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference'] + list(args),
                       capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = run({}, '--steps', '1', '--warmup', '1')
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert [k for k in NEED if k not in d] == []
    assert d['impl'] == 'reference' and d['metric'] == d['unit'] == 'Gcell-steps/s'
    assert d['steps'] == 1 and d['n_gpus'] == 1 and d['higher_is_better'] is True
    assert d['value'] > 0 and d['dtype'] == 'f32' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and '32768' in d['config']['workload']
    cb, e2e = d['cpu_baseline'], d['e2e']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['sample']
    assert cb['value'] == d['value'] == e2e['value'] and e2e['unit'] == d['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


def test_reference_arm_is_rank0_only_under_a_multi_rank_launch():
    lines = run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}, '--gpus', '2', '--steps', '1', '--warmup', '1')
    assert lines == []
