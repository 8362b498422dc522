"""Dev tool: wall-clock phases of one sharded encode + decode of the vox12 cloud on ONE rank (where the serial time goes)."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgcv1_b200 import runtime, sharding, synthetic
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29611", rank=0, world_size=1)
cubes, pos, nums = synthetic.workload("vox12", seed=0)
B = len(cubes)
lc = sharding.GpuLocalCodec("voxception", "", 0)
x = torch.from_numpy(cubes).pin_memory().to(lc.codec.dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(2):
    t0 = T()
    part = lc.encode_local_packed(x)
    t1 = T()
    packed = part.pop("y_packed")
    blob = runtime.to_host(packed, "stream_blob").copy()
    t2 = T()
    zs, zmin, zmax = lc.encode_z(part["z_hat"])
    t3 = T()
    zh = lc.decode_z(zs, zmin, zmax, np.array(part["z_hat"].shape, np.int32))
    t4 = T()
    up = lc.codec.to_device(blob)
    off = np.zeros(B + 1, np.int64); np.cumsum(part["y_lens"], out=off[1:])
    t5 = T()
    res = lc.decode_local_points(None, part["y_min"], part["y_max"], zh, nums, 1.0, uploaded=(up, lc.codec.to_device(off)))
    t6 = T()
    print("encode_local_packed %.3f | blob D2H %.3f (%d MB) | encode_z %.3f (%d B) | decode_z %.3f | blob H2D %.3f | decode_local_points %.3f | total %.3f s -> %.0f cubes/s"
          % (t1 - t0, t2 - t1, len(blob) >> 20, t3 - t2, len(zs), t4 - t3, t5 - t4, t6 - t5, t6 - t0, B / (t6 - t0)))
