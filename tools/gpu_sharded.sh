#!/bin/bash
# Config 3 (vox12 cloud sharded over the GPUs, strong scaling) and config 1 at every N up to the box's GPU count.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TAG=${1:-r02}
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N))"; fi
  timeout 600 $L bench.py --config 3 --gpus $N --steps 3 --warmup 2 > gpurun_out/${TAG}_sharded_n$N.json 2> gpurun_out/${TAG}_sharded_n$N.err
  tail -c 300 gpurun_out/${TAG}_sharded_n$N.err
  python - <<EOF
import json
try:
    s = open("gpurun_out/${TAG}_sharded_n$N.json").read()
    d = json.loads(s[s.index("{"):])            # torchrun ranks may print an NCCL banner before the line
    print("config3 N=$N value", d["value"], "e2e", d["e2e"]["value"], "sha", d["stream_sha256_16"], "bytes", d["stream_bytes"])
except Exception as e:
    print("config3 N=$N failed", e)
EOF
done
