#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -2
timeout 60 python tools/bench_conv.py 64 2>&1 | grep -E "dbg=|conv_umma"
