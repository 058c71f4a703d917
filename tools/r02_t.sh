#!/bin/bash
# Cornell/sphere kernels (no trees): 8-entry FIFO at 8 / 7 / 6 CTAs per SM, and the block-ahead buffer (BUF_NEXT) at 8 / 7
set -u
mkdir -p gpurun_out
for tag in main l7 l6 m4 m4l7; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  RPTB_LIB=$PWD/$lib timeout 600 python tools/gpu_bvh_ab.py $tag sphere cornell > gpurun_out/r02t_$tag.log 2>&1; echo "exit $?" >> gpurun_out/r02t_$tag.log
done
cat gpurun_out/r02t_*.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print(r['tag'], r['config'], round(r['Msamples_s'], 1), r['image_mean'])
    elif l.strip() and not l.startswith('exit 0'): print(l.rstrip()[:200])"
