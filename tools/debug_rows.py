"""Dev tool: where the GPU's CDF rows differ from the host twin's."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgcv1_b200 import runtime, synthetic
from pcgcv1_b200.models import conditional_entropy_model
codec = runtime.get_codec("voxception", "")
cubes, _, _n = synthetic.workload("vox10", seed=0, max_cubes=6)
x = codec.to_device(cubes)
y = codec.analysis(x)
z = torch.round(codec.hyper_encode(y))
loc, scale = codec.hyper_decode(z, 1e-9)
B = y.shape[0]
y2, l2, s2 = y.reshape(B, -1), loc.reshape(B, -1), scale.reshape(B, -1)
cem = conditional_entropy_model.SymmetricConditional().bind(codec)
iv, mm = cem.intervals_dev(y2, l2, s2)
mm_h = mm.cpu().numpy()
rows_gpu, off = codec.laplace_cdf(l2, s2, mm_h)
rows_gpu = rows_gpu.cpu().numpy().view(np.uint16)
rows_cpu, off2 = runtime.host_laplace_cdf(l2.cpu().numpy(), s2.cpu().numpy(), mm_h)
d = np.nonzero(rows_gpu != rows_cpu)[0]
print("entries", rows_cpu.size, "differ", d.size, "finite loc", bool(torch.isfinite(l2).all()), "finite scale", bool(torch.isfinite(s2).all()),
      "scale min/max", float(s2.min()), float(s2.max()), "loc min/max", float(l2.min()), float(l2.max()))
E = l2.shape[1]
lh, sh = l2.cpu().numpy(), s2.cpu().numpy()
seen = set()
for i in d[:400]:
    b = int(np.searchsorted(off, i, side="right") - 1)
    n = int(mm_h[b, 1] - mm_h[b, 0] + 1)
    e = (i - off[b]) // n
    if (b, e) in seen: continue
    seen.add((b, e))
    if len(seen) > 6: break
    print("cube", b, "elem", e, "N", n, "min", mm_h[b, 0], "loc %.9g scale %.9g" % (lh[b, e], sh[b, e]))
    print("  gpu", rows_gpu[off[b] + e * n: off[b] + (e + 1) * n])
    print("  cpu", rows_cpu[off[b] + e * n: off[b] + (e + 1) * n])
