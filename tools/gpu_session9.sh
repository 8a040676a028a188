#!/bin/bash
# session 9: GF(2^8) vector kernels, single-accumulator small recombine, party-loop unroll variants, codec
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
summ() { python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$1.json')); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
print('$1 value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f e2e %s dropin %s small %s' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac'], (d.get('e2e') or {}).get('value'), json.dumps(d.get('e2e_dropin'))[:600], d.get('small_call')))"; }
for v in libmpyc_b200.so lib_mu2.so lib_mu3.so lib_mu4.so; do
  for w in c5 c3 ns64; do
    MPYC_B200_LIB=$v timeout 300 python bench.py --workload $w --steps 10 --no-cpu --no-e2e > gpurun_out/bench_${w}_$v.json 2>>gpurun_out/variants.err; summ ${w}_$v
  done
done
for w in c3g c4; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/bench_$w.json 2>>gpurun_out/variants.err; summ $w
done
timeout 300 python bench.py --workload c3 --steps 10 --no-cpu > gpurun_out/bench_c3full.json 2>>gpurun_out/variants.err; summ c3full
timeout 300 python tools/time_inverse.py > gpurun_out/inverse.txt 2>&1; cat gpurun_out/inverse.txt
timeout 300 python tools/profile_dropin.py > gpurun_out/profile_dropin.txt 2>&1; grep -v "^$" gpurun_out/profile_dropin.txt | head -60
tail -3 gpurun_out/variants.err
