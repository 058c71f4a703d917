#!/bin/bash
# megakernel vs wavefront on the pegasus-derived dragon proxy (BVH), + ncu of the wavefront trace kernel
set -u
mkdir -p gpurun_out
for eng in megakernel wavefront; do
  timeout 600 python bench.py --workload dragon --engine $eng --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_dragon_$eng.json 2> gpurun_out/r02b_bench_dragon_$eng.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02b_mega_pegasus \
    python bench.py --workload dragon --engine megakernel --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02b_ncu_mega.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_trace_kernel -s 12 -c 1 -f -o gpurun_out/r02b_wf_trace_pegasus \
    python bench.py --workload dragon --engine wavefront --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02b_ncu_wf.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/r02b_launches_wf.csv \
    python bench.py --workload dragon --engine wavefront --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02b_launches_wf.log 2>&1
ls -la gpurun_out
