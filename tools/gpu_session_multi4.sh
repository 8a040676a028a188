#!/bin/bash
# 2-GPU session after the DST-template change: K2 regression check on one GPU, then reshare tests + bench on two
mkdir -p gpurun_out
for w in c3 c5; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2>>gpurun_out/bench.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
print('$w value %.3e split %.3f rec %.3f step %.3f' % (d['value'], r['frac'], r['recombine']['frac'], r['step_total']['frac']))"
done
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split or generate or chacha") > gpurun_out/pytest_split.log 2>&1; echo "pytest split rc=$?"; tail -2 gpurun_out/pytest_split.log
bash tools/gpu_session_multi3.sh 2
