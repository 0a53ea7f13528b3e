#!/bin/bash
# launch list of one single-worker verification run under ncu (per-launch device time and pipe utilisation)
# usage (on the GPU box): scripts/launch_list.sh <tag>
export FXG_WORKERS=1
python scripts/prof_run.py > gpurun_out/plain_$1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/launches_$1.csv python scripts/prof_run.py > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/plain_$1.log
