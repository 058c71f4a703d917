#!/bin/bash
# ncu pass over the wavefront engine's kernels on one workload: per-kernel-name time shares + metrics of the trace kernel
tag=$1; shift
ncu --clock-control none -k regex:wf_trace_kernel -s 20 -c 1 --csv --log-file gpurun_out/ncu_wf_$tag.csv \
  --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
  python tools/gpu_engines.py "$@" > gpurun_out/ncu_wf_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_wf_$tag.csv")) if len(r)>10]
for r in rows[1:]:
    print("%-90s %s" % (r[-3], r[-1]))
PY
