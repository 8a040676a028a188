#!/bin/bash
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gf256 or protocol or golden_split") > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in libmpyc_b200.so lib_mb5.so; do
  for w in c3 ns64 c5; do
    MPYC_B200_LIB=$v timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>>gpurun_out/variants.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
print('$v $w value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac']))"
  done
done
for w in c3g modmul_generic c4; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/bench_$w.json 2>>gpurun_out/variants.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
print('$w value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f small_call %s' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac'], d.get('small_call')))"
done
tail -3 gpurun_out/variants.err
