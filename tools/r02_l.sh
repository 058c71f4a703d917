#!/bin/bash
# triangle soup A/B (RPTB_NO_SOUP=1 = off) + suite
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02l_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02l_suite.log
for rep in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02l_cornell_soup_$rep.json 2> gpurun_out/r02l_cornell_soup_$rep.err
  RPTB_NO_SOUP=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02l_cornell_nosoup_$rep.json 2> gpurun_out/r02l_cornell_nosoup_$rep.err
done
cp build/obj/kernels_f32.o gpurun_out/r02l_kernels_f32.o
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02l_cornell \
    python bench.py --no-secondary --spp 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02l_ncu_cornell.log 2>&1
ls gpurun_out | grep r02l | wc -l
