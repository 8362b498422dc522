#!/bin/bash
# Config 3 at N GPUs for a list of PCGC_SHARD_Z_RATIO values (0 = count-balanced decode slices): tools/gpu_zratio.sh N r1 r2 ...
mkdir -p gpurun_out
N=$1; shift
for R in "$@"; do
  PCGC_SHARD_Z_RATIO=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --config 3 --gpus $N --steps 3 --warmup 2 > gpurun_out/zr_n${N}_$R.json 2> gpurun_out/zr_n${N}_$R.err
  python - <<EOF
import json
try:
    s = open("gpurun_out/zr_n${N}_$R.json").read()
    d = json.loads(s[s.index("{"):])
    print("config3 N=$N z_ratio=$R value", d["value"], "e2e", d["e2e"]["value"], "sha", d["stream_sha256_16"], "points", d.get("points_decoded"))
except Exception as e:
    print("config3 N=$N z_ratio=$R failed", e); print(open("gpurun_out/zr_n${N}_$R.err").read()[-600:])
EOF
done
