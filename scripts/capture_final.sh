#!/bin/bash
# round-2 closing evidence (on the GPU box): every capture after the same command has exited 0 without ncu
set -x
T=gpurun_out
# 1. issue-slot accounting of one config-2 job
python scripts/issue_run.py 3 > $T/r02f_issue_run.txt 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,sm__cycles_active.sum --clock-control none --csv \
    --log-file $T/r02f_issue_config2.csv python scripts/issue_run.py 3 > $T/r02f_issue_ncu.txt 2>&1
# 2. full capture of the dominant launch: the root-level score pass of a config-3 batch
python scripts/one_run.py config3 2000 3 > $T/r02f_plain_config3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dp_kernel<\(int\)8, \(bool\)1>' -c 2 \
    -f -o $T/r02f_prof_root python scripts/one_run.py config3 2000 3 > $T/r02f_ncu_full.log 2>&1
# 3. launch list of the bench command itself (config 2)
python bench.py --only config2 --steps 2 --warmup 3 --cpu-seconds 0.2 > $T/r02f_bench_plain.json 2> $T/r02f_bench_plain.err || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $T/r02f_launches_bench_config2.csv \
    python bench.py --only config2 --steps 2 --warmup 3 --cpu-seconds 0.2 > $T/r02f_bench_ncu.json 2> $T/r02f_bench_ncu.err
ls -la $T | tail -12
