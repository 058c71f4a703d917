#!/bin/bash
# Run on the GPU box:  tools/ncu_ext.sh r01
# ncu --set full of the F_EVERY megakernel (kd-trees over whole shapes) on examples/fractal_teapots, 800x600, 8 spp.
R=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/${R}_render_kernel_fractal_teapots \
    python bench.py --workload fractal_teapots --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_full_fractal_teapots.log 2>&1
ls -la gpurun_out | grep ${R}_render_kernel_fractal
