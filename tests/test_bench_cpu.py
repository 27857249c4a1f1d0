# coding: utf-8
"""bench.py contract pieces that run without a GPU: the reference arm (the CPU restatement timed with all host threads) and
the loud failure of the product arm when there is no CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    r = _run(['--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'wavenet_generation_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'cfg2' in d['config']['workload'] and d['gpu_launches'] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'], env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(['--steps', '1', '--warmup', '0', '--no-cpu-baseline'])
    assert r.returncode != 0
    assert 'cuda' in (r.stderr + r.stdout).lower()
