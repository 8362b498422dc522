#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma_modes.py -m gpu -q -x 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_cdf_kernel -s 2 -c 2 -o gpurun_out/prof_entropy -f python bench.py --steps 1 --warmup 1 --cubes 64 --no-cpu > gpurun_out/prof_entropy.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
