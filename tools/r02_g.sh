#!/bin/bash
# demand-driven RNG refill: suite + all configs, slot (default) and vx
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02g_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02g_suite.log
for vx in 0 1; do
  RPTB_VX=$vx timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02g_cornell_vx$vx.json 2> gpurun_out/r02g_cornell_vx$vx.err
  for wl in teapot dragon glass sphere; do
    RPTB_VX=$vx timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02g_${wl}_vx$vx.json 2> gpurun_out/r02g_${wl}_vx$vx.err
  done
done
ls gpurun_out | grep r02g | wc -l
