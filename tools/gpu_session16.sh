#!/bin/bash
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "matmul") > gpurun_out/pytest_matmul.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_matmul.log
timeout 600 python tools/time_matmul.py > gpurun_out/matmul.jsonl 2> gpurun_out/matmul.err; cat gpurun_out/matmul.jsonl; tail -3 gpurun_out/matmul.err
