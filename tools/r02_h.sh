#!/bin/bash
# lane-group BVH8 traversal (main) + RNG policy A/B (fifo / fill / demand=main) + BVH occupancy sweep
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02h_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02h_suite.log
for tag in main fifo fill; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  RPTB_LIB=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02h_cornell_$tag.json 2> gpurun_out/r02h_cornell_$tag.err
  for wl in teapot dragon glass sphere dragon_knot; do
    RPTB_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02h_${wl}_$tag.json 2> gpurun_out/r02h_${wl}_$tag.err
  done
done
for tag in bvh4 bvh5 bvh8; do
  for wl in teapot dragon dragon_knot; do
    RPTB_LIB=$PWD/rpt_b200/lib/librpt_b200_$tag.so timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02h_${wl}_$tag.json 2> gpurun_out/r02h_${wl}_$tag.err
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02h_dragon_coop \
    python bench.py --workload dragon --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02h_ncu_dragon.log 2>&1
ls gpurun_out | grep r02h | wc -l
