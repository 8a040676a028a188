"""Several processes sharing ONE GPU (what `python demo.py -M3` does with the engine installed): first uses of freshly
uploaded tables must be correct under time-slicing.  Regression test for the table-upload race of round 2."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():
    pytest.skip('no CUDA device', allow_module_level=True)

HERE = os.path.dirname(os.path.abspath(__file__))
PROGRAM = os.path.join(HERE, 'programs', 'first_use_tables.py')


@pytest.mark.parametrize('procs', [1, 4])
def test_first_use_of_tables_with_concurrent_processes(procs):
    running = [subprocess.Popen([sys.executable, PROGRAM, f'proc{i}', '25'], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
               for i in range(procs)]
    outs = [p.communicate(timeout=600)[0] for p in running]
    for i, (p, out) in enumerate(zip(running, outs)):
        assert p.returncode == 0, out[-2000:]
        assert f'proc{i} done, mismatches: 0' in out, out[-2000:]
