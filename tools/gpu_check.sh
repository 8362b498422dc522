#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 600 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 400 python bench.py --steps 3 --warmup 3 --cpu-cubes 4 2>&1 | tail -3 | tee gpurun_out/bench.log
timeout 60 python tools/bench_conv.py 64 2>&1 | tail -14
