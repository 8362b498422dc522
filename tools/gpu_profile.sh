#!/bin/bash
# r01 final profiles: ncu launch list of the bench command + --set full of the two dominant kernels (K_a16 z-banded, K_b16 tile)
cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launch_run.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"zband_kernel<.int.1, .int.2|conv_umma_kernel<.int.32, .int.2, .int.1, .int.2>" -s 6 -c 2 -o gpurun_out/prof_final -f python tools/bench_conv.py 64 > gpurun_out/prof_final.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-cubes 4 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read()); print('value',d['value'],'e2e',d['e2e']['value']); print(d['roofline'])"
