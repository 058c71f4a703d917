#!/bin/bash
# four-word Philox blocks in the f32 path: per-config table, Cornell bench, then the GPU suite
set -u
mkdir -p gpurun_out
timeout 600 python tools/gpu_scenes.py > gpurun_out/r02q_scenes.log 2>&1; echo "exit $?" >> gpurun_out/r02q_scenes.log
cp gpurun_out/scenes_table.json gpurun_out/r02q_scenes_table.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02q_suite.log
grep -h "Msamples_s" gpurun_out/r02q_scenes.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['config'], round(r['Msamples_s'], 1))"
cat gpurun_out/r02q_bench.json | cut -c1-600
tail -5 gpurun_out/r02q_suite.log
