#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PCGC_VERBOSE=1 timeout 600 python - <<'PY' 2>&1 | tail -40
import time, numpy as np, torch
from pcgcv1_b200 import synthetic, transform, runtime
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import model_voxception
cubes, pos, nums = synthetic.workload("vox10")
pinned = torch.from_numpy(cubes).pin_memory()
codec = runtime.get_codec("voxception", "")
for it in range(3):
    print("---- iteration", it)
    t0 = time.time()
    out = transform.compress_hyper(pinned, model_voxception, "")
    host = [o.numpy() for o in out]
    torch.cuda.synchronize(); t1 = time.time()
    xs = transform.decompress_hyper(*host, model_voxception, "")
    torch.cuda.synchronize(); t2 = time.time()
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
    t3 = time.time()
    print("compress %.1f ms, decompress %.1f ms, select %.1f ms -> %.1f cubes/s; ybytes %d" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, len(cubes)/(t3-t0), sum(len(s) for s in host[0])))
PY
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-cubes 4 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read()); print('value',d['value'],'e2e',d['e2e'],'cpu',d['cpu_baseline']['value'])"
