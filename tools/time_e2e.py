"""Dev tool: wall-clock split of one end-to-end step (compress_hyper / decompress_hyper / select_voxels) per coder mode."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import model_voxception

cubes, _, nums = synthetic.workload("vox10", seed=0)
codec = runtime.get_codec("voxception", "")
pinned = torch.from_numpy(cubes).pin_memory()
for mode in (os.environ.get("MODES", "gpu,host").split(",")):
    os.environ["PCGC_CODER"] = mode
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = transform.compress_hyper(pinned, model_voxception, "")
        host = [o.numpy() for o in out]
        torch.cuda.synchronize(); t1 = time.perf_counter()
        xs = transform.decompress_hyper(*host, model_voxception, "")
        torch.cuda.synchronize(); t2 = time.perf_counter()
        mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
        torch.cuda.synchronize(); t3 = time.perf_counter()
    print("%s coder: compress %.1f ms, decompress %.1f ms, select %.1f ms, total %.1f ms -> %.0f cubes/s" %
          (mode, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, len(cubes) / (t3 - t0)))

# GPU-busy accounting per phase: per-launch events of the library (sums double count kernels that overlap on two streams)
os.environ["PCGC_CODER"] = "gpu"
def prof(fn):
    codec.profile(True); codec.profile_report()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    rep = codec.profile_report(); codec.profile(False)
    coder = sum(x["ms"] for x in rep if x["tag"].startswith("range_"))
    cdf = sum(x["ms"] for x in rep if x["tag"] in ("laplace_cdf", "laplace_intervals"))
    rest = sum(x["ms"] for x in rep) - coder - cdf
    return r, dt, rest, cdf, coder
out, dt, rest, cdf, coder = prof(lambda: [o.numpy() for o in transform.compress_hyper(pinned, model_voxception, "")])
print("compress   wall %.1f ms: transform kernels %.1f ms, cdf/intervals %.1f ms, coder kernels %.1f ms" % (dt, rest, cdf, coder))
xs, dt, rest, cdf, coder = prof(lambda: transform.decompress_hyper(*out, model_voxception, ""))
print("decompress wall %.1f ms: transform kernels %.1f ms, cdf/intervals %.1f ms, coder kernels %.1f ms" % (dt, rest, cdf, coder))
