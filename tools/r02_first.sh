#!/bin/bash
# Round-2 first GPU call: state of the suite on hardware + first ncu passes of the BVH megakernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.log 2>&1
nproc > gpurun_out/r02a_nproc.log; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02a_nproc.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02a_suite.log
RPTB_EXT_BVH=1 timeout 600 python -m pytest tests/test_gpu_instancing.py -m gpu -q > gpurun_out/r02a_extbvh.log 2>&1; echo "exit $?" >> gpurun_out/r02a_extbvh.log
for wl in teapot dragon; do
  spp=8; [ $wl = teapot ] && spp=32
  timeout 600 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_$wl.json 2> gpurun_out/r02a_bench_$wl.err
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02a_bvh_$wl \
    python bench.py --workload $wl --spp $spp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02a_ncu_$wl.log 2>&1
done
ls -la gpurun_out
