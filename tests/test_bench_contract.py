"""CPU: the driver-facing contract of bench.py that can be checked without a GPU --
the reference arm prints one JSON line with the agreed keys, and the non-rank-0
processes of a multi-rank launch exit quietly."""
import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def run_bench(*args, env=None):
    return subprocess.run([sys.executable, str(ROOT / 'bench.py'), *args], capture_output=True, text=True,
                          env={**os.environ, **(env or {})}, timeout=600)


def test_reference_arm_prints_one_json_line():
    result = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-seconds', '1', '--gpus', '1')
    assert result.returncode == 0, result.stderr
    lines = [l for l in result.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['higher_is_better'] is True and line['unit'] == 'timesteps/s'
    assert line['value'] > 0 and line['steps'] == 1 and line['n_gpus'] == 1
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'timesteps/s',
                           'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['config']['states'] == 1440 and 'workload' in line['config']


def test_reference_arm_other_ranks_exit_quietly():
    result = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0', '--gpus', '2',
                       env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert result.returncode == 0 and result.stdout.strip() == ''


def test_graft_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    result = run_bench('--steps', '1', '--warmup', '0')
    assert result.returncode != 0 and 'no CUDA device' in (result.stderr + result.stdout)
