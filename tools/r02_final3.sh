#!/bin/bash
# the record run of the final build: default bench + reference arm (the suite and smoke ran on the same kernels in r02_final2.sh)
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02g_reference.json 2> gpurun_out/r02g_reference.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02g_smoke.log 2>&1
cut -c1-300 gpurun_out/r02g_bench.json; echo; tail -1 gpurun_out/r02g_smoke.log
