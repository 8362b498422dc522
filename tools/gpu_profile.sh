#!/bin/bash
# ncu passes on the bench command: (1) launch list with device times, (2) --set full on the dominant conv kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launch_run.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 0 -c 2 -o gpurun_out/prof_vrn16 -f $CMD > gpurun_out/prof_run.log 2>&1
echo "full rc=$?"
