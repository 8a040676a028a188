#!/bin/bash
rm -rf gpurun_out; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "dropin or protocol or golden or wire or gf256") > gpurun_out/pytest_dropin.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dropin.log
for w in c3 ns64 c5; do
timeout 300 python bench.py --workload $w --steps 5 --no-cpu > gpurun_out/bench_$w.json 2>>gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench_$w.json')); dr=d['e2e_dropin']
print('$w e2e %.3e dropin %.3e limbwire %.3e' % (d['e2e']['value'], dr['value'], dr['limb_wire']['value']))"
done
tail -3 gpurun_out/bench.err
