#!/bin/bash
# final captures of round 2 (one GPU): launch list of the default bench command, --set full of the dominant kernels
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
for wl in cornell dragon teapot glass; do
  spp=8; [ $wl = cornell ] && spp=32; [ $wl = teapot ] && spp=32; [ $wl = glass ] && spp=64
  extra="--workload $wl"; [ $wl = cornell ] && extra="--no-secondary"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02_final_$wl \
      python bench.py $extra --spp $spp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_final_ncu_$wl.log 2>&1
done
cp build/obj/kernels_f32.o gpurun_out/r02_final_kernels_f32.o
ls -la gpurun_out | grep r02_final
for tag in b9 b10; do
  for wl in dragon dragon_knot teapot; do
    RPTB_LIB=$PWD/rpt_b200/lib/librpt_b200_$tag.so timeout 300 python bench.py --workload $wl --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02k_${wl}_$tag.json 2> gpurun_out/r02k_${wl}_$tag.err
  done
done
timeout 900 python bench.py > gpurun_out/r02k_bench_default.json 2> gpurun_out/r02k_bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02k_bench_reference.json 2> gpurun_out/r02k_bench_reference.err
