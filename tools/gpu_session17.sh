#!/bin/bash
rm -rf gpurun_out; mkdir -p gpurun_out
timeout 600 python tools/time_small_calls.py > gpurun_out/small_calls.jsonl 2> gpurun_out/small_calls.err; cat gpurun_out/small_calls.jsonl; tail -3 gpurun_out/small_calls.err
timeout 120 python -X importtime -c "import mpyc_b200.thresha" 2>&1 | sort -t'|' -k2 -n | tail -4
