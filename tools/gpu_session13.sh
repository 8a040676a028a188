#!/bin/bash
# round-end rehearsal after the ShareDst / K4 / PeerReshare changes: smoke, all gpu tests, every bench workload, evidence refresh
rm -rf gpurun_out; mkdir -p gpurun_out
ls mpyc_b200/csrc/_obj 2>/dev/null | head -3
(timeout 300 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py > gpurun_out/bench_c3.json 2>gpurun_out/bench.err ) 2>&1 | grep real; echo "bench default rc=$?"
for w in ns64 c5 modmul c4 modmul_generic c3g; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/bench_$w.json 2>>gpurun_out/bench.err
done
timeout 600 python bench.py --workload prss --steps 10 > gpurun_out/bench_prss.json 2>>gpurun_out/bench.err
for w in c3 ns64 c5 modmul c4 modmul_generic c3g; do
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
rec=r.get('recombine',{'achieved':0,'frac':0}); st=r.get('step_total',{'frac':0})
e2e=(d.get('e2e') or {}).get('value',0); dr=(d.get('e2e_dropin') or {}); cpu=(d.get('cpu_baseline') or {}).get('value',0)
print('$w value %.3e %s %.0f GB/s (%.3f) rec %.0f GB/s (%.3f) step frac %.3f e2e %.3e dropin %.3e limbwire %.3e cpu %.3e small %s' % (d['value'], r['kernel'], r['achieved'], r['frac'], rec['achieved'], rec['frac'], st['frac'], e2e, dr.get('value',0), (dr.get('limb_wire') or {}).get('value',0), cpu, d.get('small_call')))"
done
python -c "
import json
d=json.load(open('gpurun_out/bench_prss.json')); r=d['roofline']
print('prss value %.3e %.0f GB/s (%.3f) ms %.3f e2e %.3e dropin %.3e cpu %s' % (d['value'], r['achieved'], r['frac'], r['ms'], d['e2e']['value'], d['e2e']['dropin']['value'], d['cpu_baseline']['value']))"
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>>gpurun_out/bench.err ) 2>&1 | grep real; cut -c1-200 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
cap() {  # name, kernel regex, workload args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/prof_$1 python bench.py $3 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$1.log 2>&1
}
for w in c3 ns64 c5 c3g; do cap split_$w k_split "--workload $w"; done
cap rec_c3g k_recombine "--workload c3g"
for f in gpurun_out/prof_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.csv 2>/dev/null; done
for f in gpurun_out/prof_*.ncu-rep; do case $f in *prof_split_c3.ncu-rep) ;; *) rm -f $f;; esac; done
rm -f gpurun_out/ncu_*.log
du -sh gpurun_out
tail -3 gpurun_out/bench.err
