#!/bin/bash
# 2 GPUs: the multi-device handle tests + bench.py --gpus 2 exactly as the driver launches it
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r02n2_multi.log 2>&1; echo "exit $?" >> gpurun_out/r02n2_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02n2_bench.json 2> gpurun_out/r02n2_bench.err
tail -c 600 gpurun_out/r02n2_bench.err
