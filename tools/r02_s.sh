#!/bin/bash
# Rng<float> buffering, second pass: m3 (two blocks side by side) against m2 (8-entry FIFO), m1 and the default; then the headline bench for m2 and m3
set -u
mkdir -p gpurun_out
for tag in m3 m2 m1 main; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  RPTB_LIB=$PWD/$lib timeout 600 python tools/gpu_bvh_ab.py $tag sphere cornell glass teapot monomial_glass fractal_spheres > gpurun_out/r02s_$tag.log 2>&1; echo "exit $?" >> gpurun_out/r02s_$tag.log
done
for tag in m3 m2; do
  RPTB_LIB=$PWD/rpt_b200/lib/librpt_b200_$tag.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02s_bench_$tag.json 2> gpurun_out/r02s_bench_$tag.err
done
cat gpurun_out/r02s_*.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print(r['tag'], r['config'], round(r['Msamples_s'], 1), r['image_mean'])
    elif l.strip() and not l.startswith('exit 0'): print(l.rstrip()[:200])"
for tag in m3 m2; do python -c "
import json,sys; r=json.load(open('gpurun_out/r02s_bench_$tag.json')); print('$tag bench', r['value'], r['clocks'])"; done
