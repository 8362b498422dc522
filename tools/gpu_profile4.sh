#!/bin/bash
# r01 final profiles: ncu launch list of the bench command + --set full of the two dominant kernels (K_b16, K_a16 streaming)
cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launch_run.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_umma_stream_kernel<32, 2, 1" -s 4 -c 1 -o gpurun_out/prof_kb16 -f python tools/bench_conv.py 32 > gpurun_out/prof_kb16.log 2>&1
echo "kb16 rc=$?"
ls -la gpurun_out/
