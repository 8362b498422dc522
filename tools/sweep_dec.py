"""Dev tool: decompress_hyper wall time against the decode chunk schedule (PCGC_DEC_RAMP / PCGC_DEC_CHUNK)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.models import model_voxception
from pcgcv1_b200.dataprocess import inout_points

cubes, _, nums = synthetic.workload("vox10", seed=0)
codec = runtime.get_codec("voxception", "")
pinned = torch.from_numpy(cubes).pin_memory()
out = transform.compress_hyper(pinned, model_voxception, "")
host = [o.numpy() for o in out]
ref = None
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o2 = transform.compress_hyper(pinned, model_voxception, "")
    h2 = [o.numpy() for o in o2]
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
print("compress %.2f ms; z string identical: %s" % (dt, bytes(h2[4]) == bytes(host[4])))
refm = None
SHORT = bool(int(os.environ.get("SWEEP_SHORT", "0")))
for ramp, rest in [("64", 512), ("8,24,64", 512), ("8,32,64", 512), ("24,48", 512)] if SHORT else [("64", 512), ("8,24,64", 512), ("16,32,64", 512), ("8,32,64", 512), ("16,48,64", 512), ("16,32,48", 512), ("24,48", 512), ("16,48", 512), ("32,64", 512), ("16,32,64", 64)]:
    os.environ["PCGC_DEC_RAMP"] = ramp; os.environ["PCGC_DEC_CHUNK"] = str(rest)
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        xs = transform.decompress_hyper(*host, model_voxception, "")
        mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    x = xs.tensor
    if ref is None:
        ref = x.clone(); refm = inout_points.select_voxels(x, nums, 1.0, codec=codec, dtype="uint8")
    same = bool(torch.equal(ref, x)) and np.array_equal(mask, refm)
    print("ramp %-12s rest %3d: decompress+select median %.2f ms (min %.2f) identical=%s" % (ramp, rest, sorted(ts[1:])[2], min(ts[1:]), same))
