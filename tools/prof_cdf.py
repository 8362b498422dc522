"""Dev tool: one call of the per-element CDF row kernel and of the intervals kernel on cubes of the vox10 workload (for ncu)."""
import os, sys
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
for _ in range(3):
    iv, mm = cem.intervals_dev(y.reshape(B, -1), loc.reshape(B, -1), scale.reshape(B, -1))
    mm_h = mm.cpu().numpy()
    rows, off = codec.laplace_cdf(loc.reshape(B, -1), scale.reshape(B, -1), mm_h)
torch.cuda.synchronize()
print("N per cube:", sorted(set((mm_h[:, 1] - mm_h[:, 0] + 1).tolist())))
