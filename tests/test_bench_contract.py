"""The reference arm of bench.py (CPU only) prints ONE JSON line with the keys the bench contract names; the GPU arm
refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                          env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = run('--impl', 'reference', '--steps', '1', '--warmup', '1', '--batch', '32', '--config', 'mutag_max')
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'train query-graphs/s (fwd+bwd)' and d['unit'] == 'query-graphs/s'
    for key in ('value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
                'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['value'] > 0 and d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['config']['batch_per_type'] == 32 and 'workload' in d['config']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_gpu_arm_refuses_to_run_without_a_device():
    p = run('--steps', '1', '--warmup', '1', '--batch', '32')
    assert p.returncode != 0
    assert 'no CPU fallback' in (p.stderr + p.stdout)
