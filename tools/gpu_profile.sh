#!/bin/bash
# ncu passes on a short bench run: (1) launch list with device times, (2) --set full on the dominant kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cubes 32 --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launch_run.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 12 -c 6 -o gpurun_out/prof_umma -f $CMD > gpurun_out/prof_run.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail
