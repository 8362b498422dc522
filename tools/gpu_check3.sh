#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-cubes 4 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read()); print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'cpu',d['cpu_baseline']['value']); print(d['roofline']); print([(k['tag'],k['share']) for k in d['kernels'][:8]])"
