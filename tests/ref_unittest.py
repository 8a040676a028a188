"""Program for tests/run_installed.py: run one of the REFERENCE's own unittest files (path in argv[1]) in this process,
i.e. with whatever run_installed.py installed behind mpyc.thresha / mpyc.finfields.  Exit status 0 iff all tests pass."""
import os
import sys
import unittest

path = os.path.abspath(sys.argv[1])
sys.argv = [sys.argv[0]]
suite = unittest.defaultTestLoader.discover(os.path.dirname(path), pattern=os.path.basename(path))
result = unittest.TextTestRunner(verbosity=1).run(suite)
print(f'REFTESTS run={result.testsRun} failures={len(result.failures)} errors={len(result.errors)}')
sys.exit(0 if result.wasSuccessful() and result.testsRun else 1)
