#!/bin/bash
# 8 GPUs: bench.py --gpus 8 exactly as the driver launches it (weak-scaling Cornell + strong-scaling dragon / glass), N = 4 too, + multi-device tests
set -u
mkdir -p gpurun_out
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02n8_bench_n$n.json 2> gpurun_out/r02n8_bench_n$n.err
done
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k "multi_device" > gpurun_out/r02n8_multi.log 2>&1; echo "exit $?" >> gpurun_out/r02n8_multi.log
tail -c 400 gpurun_out/r02n8_bench_n8.err
