#!/bin/bash
# multi-GPU session: N given as $1 -- 2-GPU tests (sharding + co-located reshare), scaling bench, reshare bench
N=$1
mkdir -p gpurun_out
nvidia-smi -L
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q) > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
for w in c3 ns64; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $w > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
  echo "rc=$?"; tail -2 gpurun_out/scale_${w}_n$N.err | cut -c1-300
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_${w}_n$N.json') if l.startswith('{')][-1]); r=d['roofline']
print('$w N=$N value %.3e ms/step %.3f split %.3f rec %.3f e2e %.3e launches %d clocks %s' % (d['value'], d['ms_per_step'], r['frac'], r['recombine']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks']))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_reshare.py > gpurun_out/reshare_n$N.json 2> gpurun_out/reshare_n$N.err; echo "reshare rc=$?"; cat gpurun_out/reshare_n$N.json; tail -3 gpurun_out/reshare_n$N.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/ref_n$N.json 2>gpurun_out/ref_n$N.err; cat gpurun_out/ref_n$N.json | cut -c1-300
