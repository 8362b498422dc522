"""Dev tool: device time of the interval kernel (encoder side) and of the CDF-row kernel (decoder side) on 64 cubes of the vox10
workload, CUDA events around 10 calls each.  PCGC_LIB=<variant .so> times an A/B build of entropy.cu; the interval words and rows
are hashed so that variants can be checked for identical results."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pcgcv1_b200 import runtime, synthetic
from pcgcv1_b200.models import conditional_entropy_model

codec = runtime.get_codec("voxception", "")
cubes, _, nums = synthetic.workload("vox10", seed=0, max_cubes=64)
y = codec.analysis(codec.to_device(cubes))
z_hat = torch.round(codec.hyper_encode(y))
loc, scale = codec.hyper_decode(z_hat, 1e-9)
B = y.shape[0]
cem = conditional_entropy_model.SymmetricConditional().bind(codec)
yf, lf, sf = y.reshape(B, -1), loc.reshape(B, -1), scale.reshape(B, -1)
iv, mm = cem.intervals_dev(yf, lf, sf)
mm_h = mm.cpu().numpy()
rows, off = codec.laplace_cdf(lf, sf, mm_h)
torch.cuda.synchronize()
h = hashlib.sha256(iv.cpu().numpy().tobytes()); h.update(rows.cpu().numpy().tobytes())

def timed(fn, n=10):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

t_iv = timed(lambda: cem.intervals_dev(yf, lf, sf))
t_rows = timed(lambda: codec.laplace_cdf(lf, sf, mm_h))
print("lib %s: intervals (incl. quantise) %.3f ms, cdf rows %.3f ms per 64 cubes, hash %s" % (os.path.basename(os.environ.get("PCGC_LIB", "default")), t_iv, t_rows,
                                                                                         h.hexdigest()[:16]))
