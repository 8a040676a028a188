#!/bin/bash
# session 10: PRSS pipeline (host SHAKE128 threads + tiled K4), per-L party-loop unroll default, inverse-kernel profile
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
summ() { python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$1.json')); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
print('$1 value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f e2e %s' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac'], json.dumps(d.get('e2e'))[:700]))"; }
for w in c3 c5 ns64 c3g; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2>>gpurun_out/variants.err; summ $w
done
MPYC_B200_LIB=lib_u2.so timeout 300 python bench.py --workload ns64 --steps 10 --no-cpu --no-e2e > gpurun_out/bench_ns64_u2.json 2>>gpurun_out/variants.err; summ ns64_u2
timeout 600 python bench.py --workload prss --steps 10 > gpurun_out/bench_prss.json 2>>gpurun_out/variants.err; summ prss; python -c "
import json; d=json.load(open('gpurun_out/bench_prss.json')); print('prss cpu', d['cpu_baseline'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_inv_batch|k_pow' -c 4 -o gpurun_out/ncu_inv python tools/time_inverse.py > gpurun_out/ncu_inv.log 2>&1; tail -3 gpurun_out/ncu_inv.log
ls -la gpurun_out | head -20
tail -5 gpurun_out/variants.err
