#!/bin/bash
# z-streaming kernel: parity on the UMMA shapes, then the conv micro-benchmark with and without it
cd /root/repo
PCGC_UMMA_STREAM=1 timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -6
PCGC_UMMA_STREAM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
echo "---- stream=1"; PCGC_UMMA_STREAM=1 timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13
for extra in "PCGC_STREAM_ZS=32" "PCGC_STREAM_RING=4" "PCGC_STREAM_ZS=64" "PCGC_UMMA_DBG=4" "PCGC_UMMA_DBG=1"; do
  echo "---- stream=1 $extra"; env PCGC_UMMA_STREAM=1 $extra timeout 120 python tools/bench_conv.py 64 2>&1 | tail -13 | head -6
done
