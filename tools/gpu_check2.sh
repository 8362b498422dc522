#!/bin/bash
# 2-GPU visit: full parity tests on GPU 0, then the scaling bench at N=1 and N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-cubes 4 2>&1 | tail -1 > gpurun_out/bench_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n2.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.log", "gpurun_out/bench_n2.log"):
    try:
        d = json.loads(open(f).read())
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"], "clocks", d["clocks"], "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "FAILED", e, open(f).read()[-2000:])
PY
