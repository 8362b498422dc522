#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -3
PCGC_UMMA_STREAM=1 timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -3
echo "---- stream=0"; timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13
echo "---- stream=1"; PCGC_UMMA_STREAM=1 timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13
for extra in "PCGC_STREAM_ZS=32" "PCGC_STREAM_RING=4" "PCGC_UMMA_DBG=1"; do
  echo "---- stream=1 $extra"; env PCGC_UMMA_STREAM=1 $extra timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13 | head -6
done
echo "---- stream=0 WT=2"; PCGC_UMMA_WT=2 timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13 | head -8
