#!/bin/bash
# final defaults (register FIFO, forward clamp, per-lane binary BVH without prefetch): suite + all configs; BVH CTAs/SM 6/7/8; 6-entry FIFO on Cornell
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02j_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02j_suite.log
for tag in main fifo6; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  RPTB_LIB=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02j_cornell_$tag.json 2> gpurun_out/r02j_cornell_$tag.err
  for wl in glass sphere; do
    RPTB_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02j_${wl}_$tag.json 2> gpurun_out/r02j_${wl}_$tag.err
  done
done
for tag in main b7 b8; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  for wl in teapot dragon dragon_knot; do
    RPTB_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02j_${wl}_$tag.json 2> gpurun_out/r02j_${wl}_$tag.err
  done
done
ls gpurun_out | grep r02j | wc -l
