"""Caller-level harness: the UNMODIFIED reference runtime (mpyc/runtime.py, sectypes.py, the demos and the reference's
own unittest files) running on top of mpyc_b200.install().

Two device modes, same tests:
  * 'oracle'  (build container, no GPU; not gpu-marked): the four device round trips of mpyc_b200.thresha are answered
    by tests/oracle_device.py, everything above them is the product code.  Needs the reference checkout
    (/root/reference or $MPYC_REFERENCE); skipped where it is absent.
  * 'cuda'    (-m gpu): the real kernels through the C ABI.  On the GPU box the reference is importable from
    baseline/_ref (the `mpyc` package only: pip install --target of the unmodified checkout, git-ignored) or
    $MPYC_REFERENCE; tests that need the reference's tests/ or demos/ directories skip when only the package is there.

Every run goes through tests/run_installed.py in a child process (MPyC parses sys.argv and builds its runtime at import;
-M3 re-executes the command line for the other parties, mpyc/runtime.py:5156-5189).
Pinned here: reference tests/test_thresha.py, test_finfields.py, test_runtime.py (20/20) under install(); np_aes.py
(FIPS-197 ciphertext, docs/demos.rst:606-611) and np_cnnmnist.py (logits identical to the uninstalled run), 1 party
and -M3; and tests/programs/secure_ops.py, installed == uninstalled, 1 party and -M3.
"""
import itertools
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LAUNCHER = os.path.join(HERE, 'run_installed.py')
PROGRAM = os.path.join(HERE, 'programs', 'secure_ops.py')
FIPS197 = '69c4e0d86a7b0430d8cdb78070b4c55a'


def _find_reference():
    for cand in (os.environ.get('MPYC_REFERENCE'), '/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'mpyc')):
            return cand
    return None


REF = _find_reference()
HAVE_TESTS = bool(REF) and os.path.isdir(os.path.join(REF, 'tests'))
HAVE_DEMOS = bool(REF) and os.path.isfile(os.path.join(REF, 'demos', 'np_aes.py'))
_ports = itertools.count(13000 + (os.getpid() % 400) * 40, 8)


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except ImportError:
        return False


@pytest.fixture(params=['oracle', pytest.param('cuda', marks=pytest.mark.gpu)])
def mode(request):
    if REF is None:
        pytest.skip('no reference available (MPYC_REFERENCE, /root/reference or baseline/_ref)')
    if request.param == 'cuda' and not _has_cuda():
        pytest.skip('no CUDA device')
    if request.param == 'oracle' and _has_cuda():
        pytest.skip('GPU present: the cuda variant runs the real kernels instead of the oracle stand-in')
    return request.param


def run(mode, program, args=(), flags=('install',), cwd=None, timeout=900, parties=1, env_extra=None):
    flags = list(flags)
    if mode == 'oracle' and 'off' not in flags:
        flags.append('oracle')
    env = dict(os.environ, MPYC_REFERENCE=REF, MPYC_B200_HARNESS=','.join(flags), PYTHONDONTWRITEBYTECODE='1')
    env.update(env_extra or {})
    cmd = [sys.executable, LAUNCHER, program, *args, '--no-log']
    if parties > 1:
        cmd += [f'-M{parties}', '-B', str(next(_ports))]
    r = subprocess.run(cmd, cwd=cwd or HERE, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, f'{cmd} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}'
    return r.stdout


@pytest.mark.parametrize('flags', [('install',), ('install', 'limb_wire'), ('install', 'ops')], ids=lambda f: '+'.join(f))
@pytest.mark.parametrize('name', ['test_thresha', 'test_finfields', 'test_runtime'])
def test_reference_unittests_under_install(mode, name, flags):
    """The reference's own unit tests, all green with the engine behind mpyc.thresha (and, with 'ops', behind the
    FiniteFieldArray operators and batched inverse/pow/sqrt)."""
    if not HAVE_TESTS:
        pytest.skip('reference tests/ directory not available (package-only reference)')
    out = run(mode, os.path.join(HERE, 'ref_unittest.py'), [os.path.join(REF, 'tests', name + '.py')], flags)
    assert 'failures=0 errors=0' in out, out[-2000:]
    if name == 'test_runtime':
        assert 'REFTESTS run=20 ' in out


@pytest.mark.parametrize('flags', [('install',), ('install', 'limb_wire', 'ops')], ids=lambda f: '+'.join(f))
@pytest.mark.parametrize('parties', [1, 3])
def test_np_aes_fips197(mode, parties, flags):
    """BASELINE configs[3]: demos/np_aes.py unchanged; AES-128 of the FIPS-197 example block."""
    if not HAVE_DEMOS:
        pytest.skip('reference demos/ not available (package-only reference)')
    out = run(mode, 'np_aes.py', ['-1'], flags, cwd=os.path.join(REF, 'demos'), parties=parties)
    assert f'Ciphertext:  {FIPS197}' in out, out


@pytest.mark.parametrize('parties', [1, 3])
def test_np_cnnmnist_logits_identical(mode, parties):
    """demos/np_cnnmnist.py unchanged (one image, offset 0): prediction and the 10 opened logits identical to the run
    without the engine."""
    if not HAVE_DEMOS:
        pytest.skip('reference demos/ not available (package-only reference)')
    cwd = os.path.join(REF, 'demos')
    want = run(mode, 'np_cnnmnist.py', ['1', '0'], ('off',), cwd=cwd, parties=parties)
    got = run(mode, 'np_cnnmnist.py', ['1', '0'], ('install', 'ops'), cwd=cwd, parties=parties)
    tail = lambda s: s[s.index('Image #0'):]   # noqa: E731
    assert 'with label 7: 7 predicted' in got
    assert tail(got) == tail(want)


@pytest.mark.parametrize('flags', [('install',), ('install', 'limb_wire', 'ops')], ids=lambda f: '+'.join(f))
@pytest.mark.parametrize('parties', [1, 3])
def test_secure_ops_program_installed_equals_reference(mode, parties, flags):
    """tests/programs/secure_ops.py (input, output, multiply, matmul, comparisons, random bits, fixed-point truncation,
    convert, GF(2^8) inversion, field division): opened results identical with and without the engine."""
    want = run(mode, PROGRAM, ['48'], ('off',), parties=parties)
    got = run(mode, PROGRAM, ['48'], flags, parties=parties)
    assert 'field division' in got
    assert got == want
