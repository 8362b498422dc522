#!/bin/bash
for d in 0 1 2 3 4 7; do
  PCGC_UMMA_DBG=$d timeout 300 python tools/bench_conv.py 64 2>&1 | tail -14
done
