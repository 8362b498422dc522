#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
echo "---- default"; timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13 | head -8
echo "---- dbg=1"; PCGC_UMMA_DBG=1 timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13 | head -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_stream_kernel -s 12 -c 2 -o gpurun_out/prof_stream -f python tools/bench_conv.py 32 > gpurun_out/prof_stream.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
