#!/bin/bash
# last check of round 2 on one GPU: the whole GPU suite, smoke(), the default bench line and the reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02z_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02z_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02z_smoke.log
timeout 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02z_reference.json 2> gpurun_out/r02z_reference.err
for wl in sphere teapot; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_$wl.json 2> gpurun_out/r02z_$wl.err
done
