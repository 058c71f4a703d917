#!/bin/bash
# how Rng<float> buffers the four draws of a block: main (two + two waiting), m1 (fill when empty), m2 (8-entry FIFO), ring, ringfill
set -u
mkdir -p gpurun_out
for tag in main m1 m2 ring ringfill; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  RPTB_LIB=$PWD/$lib timeout 600 python tools/gpu_bvh_ab.py $tag sphere cornell glass teapot > gpurun_out/r02r_$tag.log 2>&1; echo "exit $?" >> gpurun_out/r02r_$tag.log
done
cat gpurun_out/r02r_*.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print(r['tag'], r['config'], round(r['Msamples_s'], 1), r['image_mean'])
    elif l.strip() and not l.startswith('exit 0'): print(l.rstrip()[:200])"
