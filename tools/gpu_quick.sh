#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_umma.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/bench_conv.py 64 2>&1 | tail -14
