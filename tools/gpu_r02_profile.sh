#!/bin/bash
# r02 measurement bundle (one gpurun call): bench lines of every BASELINE config at N=1, the ncu launch list of the default bench
# command, and one --set full capture of the two dominant conv kernels.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_config1.json 2> gpurun_out/r02_bench_config1.err
python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_config2.json 2> gpurun_out/r02_bench_config2.err
python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r02_bench_config3_n1.json 2> gpurun_out/r02_bench_config3_n1.err
python bench.py --config 4 --steps 20 > gpurun_out/r02_bench_config4.json 2> gpurun_out/r02_bench_config4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu > gpurun_out/r02_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv_umma_kernel|conv_umma_zband_kernel' -s 40 -c 6 -o gpurun_out/r02_prof_conv python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu > gpurun_out/r02_prof_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'range_' -c 3 -o gpurun_out/r02_prof_coder python tools/bench_coder.py 16 > gpurun_out/r02_prof_coder.log 2>&1
tail -c 400 gpurun_out/r02_bench_config3_n1.json; tail -2 gpurun_out/r02_bench_config3_n1.err
