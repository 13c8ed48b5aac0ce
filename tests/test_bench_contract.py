"""bench.py contract checks that need no GPU: the reference arm's JSON line (timed here on the CPU oracle port), rank
gating under torchrun-style environments, and the product arm failing loudly without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.pop('RANK', None)
    e.pop('WORLD_SIZE', None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    r = run(['--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d['impl'] == 'reference'
    assert d['metric'] == 'policy fwd passes/sec at 100^3 voxels' and d['unit'] == 'passes/s'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['steps'] == 1 and d['n_gpus'] == 1 and d['value'] > 0 and d['ms_per_step'] > 0
    assert 'workload' in d['config'] and '100^3' in d['config']['workload']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['sample'] and cb['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'passes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_are_silent():
    r = run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return                                   # on a GPU box this is the real bench; covered by the driver
    r = run(['--steps', '1', '--warmup', '0'], timeout=300)
    assert r.returncode != 0                     # no silent CPU fallback
    assert not any(l.strip().startswith('{') for l in r.stdout.splitlines())
