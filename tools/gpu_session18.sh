#!/bin/bash
# last validation of the final tree: PRSS (both small forms), drop-in paths, smoke
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
(timeout 900 python -m pytest tests -m gpu -q -x -k "prss or PRSS or dropin or protocol or wire or golden_split") > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_last.log
timeout 300 python bench.py --workload prss --steps 10 --no-cpu > gpurun_out/bench_prss.json 2>gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench_prss.json')); r=d['roofline']
print('prss value %.3e %.0f GB/s (%.3f) ms %.3f e2e %.3e dropin %.3e' % (d['value'], r['achieved'], r['frac'], r['ms'], d['e2e']['value'], d['e2e']['dropin']['value']))"
tail -2 gpurun_out/bench.err
