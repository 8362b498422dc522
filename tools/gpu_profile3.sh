#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cubes 32 --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:laplace_cdf_kernel -s 2 -c 2 -o gpurun_out/prof_cdf -f $CMD > gpurun_out/prof_run3.log 2>&1
echo "full rc=$?"
