#!/bin/bash
# I/O rows: GPU tests + stage timings of the CLI on the vox10 cloud
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_io.py -m gpu -x -q 2>&1 | tail -15
mkdir -p /tmp/cli && cd /tmp/cli
timeout 600 python - <<'PY' 2>&1 | tail -60
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from pcgcv1_b200 import synthetic, test as cli
from pcgcv1_b200.dataprocess import inout_points
pts = synthetic.cloud_vox10()
inout_points.write_ply_data("vox10_cloud.ply", pts)
for it in range(2):
    t0 = time.time(); cli.main(["compress", "vox10_cloud.ply"]); t1 = time.time()
    cli.main(["decompress", "compressed/vox10_cloud"]); t2 = time.time()
    print("CLI iteration %d: compress %.3f s, decompress %.3f s" % (it, t1 - t0, t2 - t1))
rec = inout_points.load_ply_data("vox10_cloud_rec.ply")
print("input points", len(pts), "reconstructed", len(rec))
PY
