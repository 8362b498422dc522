#!/bin/bash
# last measurement bundle of round 2 (one gpurun call): bench lines (config 1, config 5 with both distortions), the ncu launch list of
# the default bench command and a --set full capture of the two CDF kernels after the water-filling threshold change.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/s10_bench_c5_bce.json 2> gpurun_out/s10_bench_c5_bce.err
python bench.py --config 5 --distortion focal --steps 5 --warmup 3 > gpurun_out/s10_bench_c5_focal.json 2> gpurun_out/s10_bench_c5_focal.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu > gpurun_out/s10_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'laplace_cdf_kernel' -s 2 -c 2 -o gpurun_out/s10_prof_cdf python tools/prof_cdf.py > gpurun_out/s10_prof_cdf.log 2>&1
tail -c 300 gpurun_out/s10_bench.err; cut -c1-200 gpurun_out/s10_bench.json; cut -c1-300 gpurun_out/s10_bench_c5_focal.json; tail -3 gpurun_out/s10_prof_cdf.log
