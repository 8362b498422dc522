#!/bin/bash
# far-field build: ncu launch list of the bench command + --set full of the tile kernel's K_b16 launches (analysis launches copy far-field tiles)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s16_launches.csv python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu > gpurun_out/s16_launch_run.log 2>&1
ncu --set full --clock-control none -k regex:'conv_umma_kernel<32, 2, 1, 2>|conv_umma_kernelILi32ELi2ELi1ELi2E' -s 9 -c 8 -o gpurun_out/s16_prof_kb16 python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu > gpurun_out/s16_prof_kb16.log 2>&1
tail -2 gpurun_out/s16_prof_kb16.log
