#!/bin/bash
# Config 1 (weak scaling, one vox10 cloud per GPU) and config 3 (one vox12 cloud sharded over the GPUs) at N = $1 GPUs -> gpurun_out/${2}_*.json
N=$1; TAG=${2:-r02}
mkdir -p gpurun_out
if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N))"; fi
$L bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
$L bench.py --config 3 --gpus $N --steps 3 --warmup 2 > gpurun_out/${TAG}_sharded_n$N.json 2> gpurun_out/${TAG}_sharded_n$N.err
python - <<EOF
import json
for f in ("bench", "sharded"):
    try:
        s = open("gpurun_out/${TAG}_%s_n$N.json" % f).read()
        d = json.loads(s[s.index("{"):])
        print(f, "N=$N value", d["value"], "e2e", d["e2e"]["value"], d.get("stream_sha256_16", ""), d.get("clocks", ""))
    except Exception as e:
        print(f, "N=$N failed", e)
EOF
