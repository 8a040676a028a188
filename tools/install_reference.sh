#!/bin/bash
# Installs the UNMODIFIED reference (lschoe/mpyc) into baseline/_ref (git-ignored; travels to the GPU box with gpurun):
#   baseline/_ref/mpyc/                 pip install --target of a copy of the read-only checkout
#   baseline/_ref/_checkout/{demos,tests}   the checkout's demo programs and unit tests, byte for byte
# Used by: bench.py --impl reference (cpu_baseline.kind "reference"), tests/test_reference_runtime.py (-m gpu: the
# reference's own runtime, tests and demos on top of the real kernels).  Nothing under baseline/ is product source.
set -e
cd "$(dirname "$0")/.."
SRC=${MPYC_REFERENCE_SRC:-/root/reference}
rm -rf /tmp/_mpyc_refcopy baseline/_ref
cp -r "$SRC" /tmp/_mpyc_refcopy
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /tmp/_mpyc_refcopy
mkdir -p baseline/_ref/_checkout
cp -r "$SRC/demos" "$SRC/tests" baseline/_ref/_checkout/
find baseline/_ref -name __pycache__ -prune -exec rm -rf {} +
rm -rf /tmp/_mpyc_refcopy
