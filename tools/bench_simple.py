"""Dev tool: per-kernel time of the model_simple transforms (config 2) with the library's per-launch events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcgcv1_b200 import runtime, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
codec = runtime.get_codec("simple", "")
cubes, _ = synthetic.surface_cubes(8, seed=3)
x = codec.to_device(np.tile(cubes, (B // 8, 1, 1, 1, 1)))
y = torch.randn(B, 8, 8, 8, 32, device=codec.dev) * 3
for _ in range(2):
    codec.analysis(x); codec.synthesis(y)
codec.profile(True); codec.profile_report()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    codec.analysis(x); codec.synthesis(y)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
prof = codec.profile_report()
prof.sort(key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in prof)
print("model_simple B=%d: %.2f ms per (analysis+synthesis) -> %.1f cubes/s, %.1f TFLOP/s" % (B, ms, B / ms * 1e3, B * 5.4169 / ms))
for r in prof[:12]:
    print("   %-40s %5.1f%%  %8.3f ms/launch  %7.2f %s" % (r["tag"], 100 * r["ms"] / tot, r["ms"] / r["count"],
          (r["flops"] / 1e12 if r["flops"] else r["bytes"] / 1e9) / (r["ms"] * 1e-3), "TF/s" if r["flops"] else "GB/s"))
