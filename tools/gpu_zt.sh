#!/bin/bash
for z in 8 4 2 1; do echo "== ZT<=$z"; PCGC_UMMA_ZT=$z timeout 300 python tools/bench_conv.py 64 2>&1 | tail -13 | head -8; done
