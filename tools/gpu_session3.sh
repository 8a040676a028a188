#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in c3 ns64 c5 modmul; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/bench_$w.json 2>>gpurun_out/bench.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
print('$w value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f e2e %.3e' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac'], d['e2e']['value']))"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_split -s 3 -c 1 -o gpurun_out/prof_split_c3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_split.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_recombine -s 3 -c 1 -o gpurun_out/prof_rec_c3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_recombine -s 3 -c 1 -o gpurun_out/prof_rec_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rec5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_split -s 3 -c 1 -o gpurun_out/prof_split_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_split5.log 2>&1
tail -3 gpurun_out/bench.err
