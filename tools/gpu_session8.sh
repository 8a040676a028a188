#!/bin/bash
# session 8: Barrett generic path, batched inverse, limb wire -- parity first, then the affected workloads
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
summ() { python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$1.json')); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
print('$1 value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f e2e %s dropin %s' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac'], (d.get('e2e') or {}).get('value'), json.dumps(d.get('e2e_dropin'))[:400]))"; }
for w in modmul_generic c3g; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/bench_$w.json 2>>gpurun_out/variants.err; summ $w
done
for w in c3 ns64 c5 modmul; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2>>gpurun_out/variants.err; summ $w
done
timeout 300 python bench.py --workload c5 --steps 5 --no-cpu > gpurun_out/bench_c5e.json 2>>gpurun_out/variants.err; summ c5e
# batched inverse vs per-element Fermat (in place) timing
timeout 300 python tools/time_inverse.py > gpurun_out/inverse.txt 2>&1; cat gpurun_out/inverse.txt
tail -3 gpurun_out/variants.err
