#!/bin/bash
# 2-GPU session: reshare tests (NCCL and peer-store modes) + reshare bench
N=$1
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x) > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -25 gpurun_out/pytest_multi.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_reshare.py > gpurun_out/reshare_n$N.json 2> gpurun_out/reshare_n$N.err; echo "reshare rc=$?"; cat gpurun_out/reshare_n$N.json; tail -8 gpurun_out/reshare_n$N.err | cut -c1-400
