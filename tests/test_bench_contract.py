"""The JSON-line contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`: the
unmodified reference where it is importable -- baseline/_ref or $MPYC_REFERENCE --, else the oracle port) prints one line
with the keys the driver reads, on the same `config` as the GPU arm, with an `e2e` that moves no bytes over PCIe."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'impl'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['unit'] == 'pairs/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['steps'] == 1 and line['n_gpus'] == 1 and line['vs_baseline'] is None
    cfg = line['config']
    assert 'workload' in cfg and cfg['p_bits'] == 128 and cfg['m'] == 5 and cfg['t'] == 2 and cfg['recombine_k'] == 3
    cb = line['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == line['value'] and cb['sample']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0
