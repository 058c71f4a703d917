#!/bin/bash
# register-FIFO RNG back as default; child prefetch, occupancy and the lane-group hybrid on the mesh configs
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02i_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02i_suite.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02i_cornell_main.json 2> gpurun_out/r02i_cornell_main.err
for wl in glass sphere; do
  timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02i_${wl}_main.json 2> gpurun_out/r02i_${wl}_main.err
done
for tag in main nopf b8 b8nopf coop4 coop8; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  for wl in teapot dragon dragon_knot; do
    RPTB_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02i_${wl}_$tag.json 2> gpurun_out/r02i_${wl}_$tag.err
  done
done
ls gpurun_out | grep r02i | wc -l
