#!/bin/bash
# GPU session: tests, unroll-variant comparison, ncu captures (C3 workload)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in lib_u1.so libmpyc_b200.so lib_u4.so; do
  for w in c3 ns64 c5; do
    echo "== $v $w"
    MPYC_B200_LIB=$v timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>>gpurun_out/variants.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e split %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f clocks %s' % (d['value'], r['achieved'], r['frac'], r['recombine']['achieved'], r['recombine']['frac'], r['step_total']['frac'], d['clocks']))"
  done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_split -s 3 -c 1 -o gpurun_out/prof_split_c3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_split.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_recombine -s 3 -c 1 -o gpurun_out/prof_rec_c3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_recombine -s 3 -c 1 -o gpurun_out/prof_rec_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rec5.log 2>&1
ls -la gpurun_out
