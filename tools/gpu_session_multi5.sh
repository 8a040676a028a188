#!/bin/bash
# N-GPU session: scaling bench (c3, ns64) + reshare bench at N ranks; ns64 single-GPU check of the 96-bit fold
N=$1
mkdir -p gpurun_out
timeout 300 python bench.py --workload ns64 --steps 10 --no-cpu --no-e2e > gpurun_out/bench_ns64.json 2>>gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench_ns64.json')); r=d['roofline']
print('ns64 1 GPU value %.3e split %.3f rec %.3f step %.3f' % (d['value'], r['frac'], r['recombine']['frac'], r['step_total']['frac']))"
for w in c3 ns64; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $w > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
  echo "rc=$?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_${w}_n$N.json') if l.startswith('{')][-1]); r=d['roofline']
print('$w N=$N value %.3e ms/step %.3f split %.3f rec %.3f e2e %.3e launches %d clocks %s' % (d['value'], d['ms_per_step'], r['frac'], r['recombine']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks']))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_reshare.py > gpurun_out/reshare_n$N.json 2> gpurun_out/reshare_n$N.err; echo "reshare rc=$?"; cat gpurun_out/reshare_n$N.json; tail -4 gpurun_out/reshare_n$N.err | cut -c1-300
