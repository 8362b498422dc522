#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for ch in 64 32 48 96; do
  PCGC_CHUNK=$ch timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_chunk$ch.log
  python -c "
import json; d=json.loads(open('gpurun_out/bench_chunk$ch.log').read()); print('chunk $ch value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])"
done
PCGC_UMMA_STREAM=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stream=2 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])"
