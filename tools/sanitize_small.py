"""Dev tool: a 3-cube compress_hyper / decompress_hyper / select_voxels round trip (the workload for compute-sanitizer runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import model_voxception
cubes, nums = synthetic.surface_cubes(3, seed=5)
codec = runtime.get_codec("voxception", "")
out = transform.compress_hyper(cubes, model_voxception, "", decompress=True)
xs = transform.decompress_hyper(*[o.numpy() for o in out[:8]], model_voxception, "")
mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
assert np.array_equal(xs.numpy(), out[8].numpy())
print("round trip ok, mask counts", mask.reshape(3, -1).sum(1), "strings", [len(s) for s in out[0].numpy()])
