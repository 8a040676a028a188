#!/bin/bash
# full round-end style session: all gpu tests, smoke, all bench workloads (with e2e, cpu baseline), ncu captures
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c3.json 2>gpurun_out/bench.err; echo "bench default rc=$?"
for w in ns64 c5 modmul c4 modmul_generic c3g; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/bench_$w.json 2>>gpurun_out/bench.err
done
for w in c3 ns64 c5 modmul c4 modmul_generic c3g; do
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
e2e=(d.get('e2e') or {}).get('value',0); dr=(d.get('e2e_dropin') or {}).get('value',0); cpu=(d.get('cpu_baseline') or {}).get('value',0)
print('$w value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f e2e %.3e dropin %.3e cpu %.3e small %s' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac'], e2e, dr, cpu, d.get('small_call')))"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>>gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
for w in c3 ns64 c5; do
  ncu --set full --clock-control none --import-source on -k regex:k_split -s 3 -c 1 -o gpurun_out/prof_split_$w python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_split_$w.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_recombine -s 3 -c 1 -o gpurun_out/prof_rec_$w python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rec_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_binop -s 3 -c 1 -o gpurun_out/prof_binop_modmul python bench.py --workload modmul --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_binop.log 2>&1
# keep the transfer small: raw-metric CSVs for every capture, the .ncu-rep only for the dominant kernel
for f in gpurun_out/prof_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.csv 2>/dev/null; done
for f in gpurun_out/prof_*.ncu-rep; do case $f in *prof_split_c3*) ;; *) rm -f $f;; esac; done
rm -f gpurun_out/ncu_*.log
du -sh gpurun_out
tail -3 gpurun_out/bench.err
