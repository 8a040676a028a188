#!/bin/bash
# session 12: K4 rewrite (wide fold) -- PRSS parity, bench, ncu capture
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "prss or PRSS or golden") > gpurun_out/pytest_prss.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_prss.log
timeout 600 python bench.py --workload prss --steps 10 > gpurun_out/bench_prss.json 2>gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench_prss.json')); r=d['roofline']
print('prss value %.3e %.0f GB/s (%.3f) ms %.3f e2e %s cpu %s' % (d['value'], r['achieved'], r['frac'], r['ms'], json.dumps(d['e2e']), d['cpu_baseline']))"
ncu --set full --clock-control none --import-source on -k regex:k_prss -s 3 -c 1 -o gpurun_out/prof_prss_prss python bench.py --workload prss --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_prss.log 2>&1
ncu -i gpurun_out/prof_prss_prss.ncu-rep --page raw --csv > gpurun_out/prof_prss_prss.csv 2>/dev/null; rm -f gpurun_out/prof_prss_prss.ncu-rep
tail -3 gpurun_out/bench.err
