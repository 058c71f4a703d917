#!/bin/bash
# Round-2 GPU call: the vertex-at-once engine (default) vs the slot engine (RPTB_VX=0): suite, benches, captures
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02e_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02e_suite.log
for vx in 1 0; do
  RPTB_VX=$vx timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02e_cornell_vx$vx.json 2> gpurun_out/r02e_cornell_vx$vx.err
  for wl in teapot dragon glass sphere dragon_knot; do
    RPTB_VX=$vx timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02e_${wl}_vx$vx.json 2> gpurun_out/r02e_${wl}_vx$vx.err
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02e_cornell_vx \
    python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r02e_ncu_cornell_vx.log 2>&1
RPTB_VX=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02e_cornell_slot \
    python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r02e_ncu_cornell_slot.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02e_dragon_vx \
    python bench.py --workload dragon --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02e_ncu_dragon_vx.log 2>&1
ls -la gpurun_out | tail -40
