#!/bin/bash
# Round-2 GPU call 4: suite with ring RNG + forward clamp composition, Cornell bench + capture, occupancy sweeps
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02d_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02d_suite.log
timeout 600 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
for tag in lite5 lite6; do
  RPTB_LIB=$PWD/rpt_b200/lib/librpt_b200_$tag.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r02d_sweep_cornell_$tag.json 2> gpurun_out/r02d_sweep_cornell_$tag.err
done
for tag in main bvh4 bvh5 bvh8; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  for wl in teapot dragon; do
    RPTB_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02d_sweep_${wl}_$tag.json 2> gpurun_out/r02d_sweep_${wl}_$tag.err
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02d_cornell \
    python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02d_ncu_cornell.log 2>&1
cp build/obj/kernels_f32.o gpurun_out/r02d_kernels_f32.o
ls -la gpurun_out | tail -30
