#!/bin/bash
# quick ncu pass on the render kernel: a handful of metrics instead of --set full
# usage (on the GPU box): tools/ncu_quick.sh <tag> [bench args...]
tag=$1; shift
ncu --clock-control none -k regex:render_kernel -s 1 -c 1 --csv --log-file gpurun_out/ncu_quick_$tag.csv \
  --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_quick_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_quick_$tag.csv")) if len(r)>10]
for r in rows[1:]:
    print("%-90s %s" % (r[-3], r[-1]))
PY
