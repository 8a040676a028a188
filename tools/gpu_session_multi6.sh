#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err
echo "rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_c3_n$N.json') if l.startswith('{')][-1]); r=d['roofline']
print('c3 N=$N value %.3e ms/step %.3f split %.3f rec %.3f e2e %.3e launches %d clocks %s' % (d['value'], d['ms_per_step'], r['frac'], r['recombine']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_reshare.py > gpurun_out/reshare_n$N.json 2> gpurun_out/reshare_n$N.err; echo "reshare rc=$?"; cat gpurun_out/reshare_n$N.json; tail -3 gpurun_out/reshare_n$N.err | cut -c1-200
