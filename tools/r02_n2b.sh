#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/gpu_e2e_multi.py 1 2 > gpurun_out/e2e_multi2.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary > gpurun_out/r02n2b_bench.json 2> gpurun_out/r02n2b_bench.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02n2b_multi.log 2>&1
