#!/bin/bash
rm -rf gpurun_out; mkdir -p gpurun_out
for v in libmpyc_b200.so lib_mb5.so lib_mb6.so; do
  for w in ns64 c3g; do
    MPYC_B200_LIB=$v timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>>gpurun_out/variants.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
print('$v $w value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac']))"
  done
done
tail -3 gpurun_out/variants.err
