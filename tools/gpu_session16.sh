#!/bin/bash
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "matmul") > gpurun_out/pytest_matmul.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_matmul.log
