#!/bin/bash
# generator moved to the next sample at the converged point (integrator.cuh, rng_ready): before (pre) / after (main), then the parity subset
set -u
mkdir -p gpurun_out
for tag in pre main; do
  lib=rpt_b200/lib/librpt_b200_$tag.so; [ $tag = main ] && lib=rpt_b200/lib/librpt_b200.so
  RPTB_LIB=$PWD/$lib timeout 300 python tools/gpu_bvh_ab.py $tag cornell sphere teapot glass > gpurun_out/r02u_$tag.log 2>&1; echo "exit $?" >> gpurun_out/r02u_$tag.log
done
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x -m gpu > gpurun_out/r02u_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02u_suite.log
cat gpurun_out/r02u_pre.log gpurun_out/r02u_main.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print(r['tag'], r['config'], round(r['Msamples_s'], 1), r['image_mean'])
    elif l.strip() and not l.startswith('exit 0'): print(l.rstrip()[:200])"
python -c "
import json; r=json.load(open('gpurun_out/r02u_bench.json')); print('bench', r['value'])"
tail -2 gpurun_out/r02u_suite.log
