"""install(): put the B200 engine behind an imported MPyC without touching MPyC's code.

mpyc/runtime.py looks the secret-sharing functions up as attributes of the `mpyc.thresha` module at
call time (runtime.py:478-485,565-572,647-656,963,982,1280,4057,4099,4219,4248); assigning the drop-in
functions of mpyc_b200.thresha over those attributes makes `_distribute`, `output`, `_reshare` and the
PRSS helpers run on the GPU while runtime.py, sectypes.py and the demos stay byte-for-byte unchanged.

Fields the engine does not cover (GF(2), GF(p^n) with n > 1, binary fields other than GF(2^8), primes
wider than 256 bits) are NOT intercepted: for those the wrapper hands the call to the reference's own
function, exactly as if install() had not been called (strict=True raises UnsupportedFieldError
instead).  For covered fields there is no fallback of any kind: without a CUDA device the call raises.
"""
import functools

from mpyc_b200 import thresha as engine
from mpyc_b200._cabi import UnsupportedFieldError
from mpyc_b200.field import context_of_field

_NAMES = ('random_split', 'recombine', 'np_random_split', 'np_recombine',
          'pseudorandom_share', 'pseudorandom_share_zero', 'np_pseudorandom_share', 'np_pseudorandom_share_0',
          '_recombination_vector', '_f_S_i')
_saved = {}


def _covered(field):
    try:
        context_of_field(field)
        return True
    except UnsupportedFieldError:
        return False


def _wrap(name, ours, theirs, strict):
    @functools.wraps(theirs)
    def call(field, *args, **kwargs):
        if _covered(field):
            return ours(field, *args, **kwargs)
        if strict:
            raise UnsupportedFieldError(f'mpyc_b200 does not cover field {getattr(field, "__name__", field)}')
        return theirs(field, *args, **kwargs)
    call.__mpyc_b200__ = True
    return call


def install(thresha_module=None, strict=False, device=0):
    """Patch `mpyc.thresha` (or the module passed in).  Returns the list of patched names."""
    if thresha_module is None:
        import mpyc.thresha as thresha_module
    if _saved:
        uninstall()
    engine.device = device
    for name in _NAMES:
        theirs = getattr(thresha_module, name)
        _saved[name] = (thresha_module, theirs)
        if name in ('_recombination_vector', '_f_S_i'):
            # keep the reference's functools.cache behaviour: results are cached per argument tuple
            ours = functools.cache(getattr(engine, name))
        else:
            ours = getattr(engine, name)
        setattr(thresha_module, name, _wrap(name, ours, theirs, strict))
    return list(_NAMES)


def uninstall():
    for name, (module, theirs) in list(_saved.items()):
        setattr(module, name, theirs)
    _saved.clear()
