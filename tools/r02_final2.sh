#!/bin/bash
# final run of round 2 after the four-word Philox change: GPU suite, smoke, default bench + reference arm, then the ncu captures
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r02f_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02f_reference.json 2> gpurun_out/r02f_reference.err
timeout 600 python tools/gpu_scenes.py > gpurun_out/r02f_scenes.log 2>&1; cp gpurun_out/scenes_table.json gpurun_out/r02f_scenes_table.json
timeout 600 python tools/gpu_scenes.py extra > gpurun_out/r02f_scenes_extra.log 2>&1; cp gpurun_out/scenes_table_extra.json gpurun_out/r02f_scenes_table_extra.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02f_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02f_launches_bench.log 2>&1
for wl in cornell dragon teapot glass; do
  spp=8; [ $wl = cornell ] && spp=32; [ $wl = teapot ] && spp=32; [ $wl = glass ] && spp=64
  extra="--workload $wl"; [ $wl = cornell ] && extra="--no-secondary"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02f_$wl \
      python bench.py $extra --spp $spp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02f_ncu_$wl.log 2>&1
done
cp build/obj/kernels_f32.o gpurun_out/r02f_kernels_f32.o
tail -3 gpurun_out/r02f_suite.log; cat gpurun_out/r02f_smoke.log | tail -1
cut -c1-400 gpurun_out/r02f_bench.json; echo; cut -c1-300 gpurun_out/r02f_reference.json; echo
grep -h Msamples_s gpurun_out/r02f_scenes.log gpurun_out/r02f_scenes_extra.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['config'], r['spp'], round(r['Msamples_s'], 1))"
ls -la gpurun_out | grep r02f_ | wc -l
