#!/bin/bash
for sb in 32 16 8 4 2; do echo "== SUB_BATCH $sb"; PCGC_SUB_BATCH=$sb timeout 300 python tools/bench_conv.py 64 2>&1 | tail -13 | head -7; done
