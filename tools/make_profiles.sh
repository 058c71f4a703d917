#!/bin/bash
# Run on the GPU box (one GPU):  tools/make_profiles.sh r01
# Produces under gpurun_out/: the ncu launch list of the default bench command, a --set full capture of
# the dominant kernel of the bench workload (Cornell, megakernel) and of the wavefront trace kernel on the
# dragon proxy.  tools/summarize_profiles.py turns them into the committed profiles/<round>_*.
set -u
R=${1:-r01}
mkdir -p gpurun_out
# (1) every launch with its device time, same command as the bench (fewer steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches_cornell.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_launches_cornell.log 2>&1
# (2) the dominant kernel, full set (reduced spp: ncu replays the launch ~40 times)
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/${R}_render_kernel_cornell \
    python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu-baseline --no-e2e > gpurun_out/${R}_full_cornell.log 2>&1
# (3) the wavefront engine on the dragon proxy: launch list + the trace kernel
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${R}_launches_dragon.csv \
    python bench.py --workload dragon --spp 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_launches_dragon.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_trace_kernel -s 20 -c 1 -f -o gpurun_out/${R}_wf_trace_dragon \
    python bench.py --workload dragon --spp 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_full_dragon.log 2>&1
ls -la gpurun_out | grep ${R}_
