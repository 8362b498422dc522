#!/bin/bash
# N = 8 on one box: the headline bench (config 1, weak scaling) and the sharded vox12 cloud (config 3, strong scaling).  Outputs -> gpurun_out/.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TAG=${1:-r02}
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $L bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$NG.json 2> gpurun_out/${TAG}_bench_n$NG.err
tail -c 300 gpurun_out/${TAG}_bench_n$NG.err
timeout 600 $L bench.py --config 3 --gpus $NG --steps 3 --warmup 2 > gpurun_out/${TAG}_sharded_n$NG.json 2> gpurun_out/${TAG}_sharded_n$NG.err
python - <<EOF
import json
for f in ("gpurun_out/${TAG}_bench_n$NG.json", "gpurun_out/${TAG}_sharded_n$NG.json"):
    try:
        s = open(f).read(); d = json.loads(s[s.index("{"):])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], d.get("stream_sha256_16", ""))
    except Exception as e:
        print(f, "failed", e)
EOF
