"""CPU test of bench.py's contract: the reference arm (the unmodified reference from baseline/_ref or /root/reference when one is
installed, else the oracle port - on the host cores) prints ONE JSON line with the keys the driver reads, and imports nothing of the
product.  No GPU involved; a single bounded sample keeps it to a few seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['unit'] == 'frame*NFE/s'
    for k in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype', 'data', 'config', 'e2e',
              'cpu_baseline'):
        assert k in d, k
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and d['value'] > 0


def test_reference_arm_does_not_import_the_product():
    """The CPU arm must not load libb2s.so (`reference.native_so_loaded` listed it): importing the product package in
    the arm's process is an error."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'];"
            "runpy.run_path('bench.py', run_name='__main__');"
            "bad = [m for m in sys.modules if m.startswith('xiaoicesing_io_b200')]; assert not bad, bad")
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
