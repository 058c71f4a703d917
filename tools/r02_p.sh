#!/bin/bash
# four-wide BVH A/B: main (BVH4, 8 CTAs/SM), bvh4_6 (6 CTAs/SM), bin (binary tree, 8 CTAs/SM)
set -u
mkdir -p gpurun_out
for tag in bin main bvh4_6 bvh4_5; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  RPTB_LIB=$PWD/$lib timeout 900 python tools/gpu_bvh_ab.py $tag > gpurun_out/r02p_$tag.log 2>&1; echo "exit $?" >> gpurun_out/r02p_$tag.log
done
timeout 600 python -m pytest tests -m gpu -q -x -k "bvh or teapot or mesh or golden or pegasus" > gpurun_out/r02p_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02p_suite.log
cat gpurun_out/r02p_*.log | grep -v "^$" | tail -40
