"""Dev tool: why the GPU encoder launch takes longer inside compress_hyper than alone."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.models.conditional_entropy_model import SymmetricConditional
codec = runtime.get_codec("voxception", "")
cubes, _, nums = synthetic.workload("vox10", seed=0)
B = len(cubes)
x = codec.to_device(cubes)
eb = transform._bottleneck(codec, 8)
cem = SymmetricConditional().bind(codec)
iv, mm_all, z_hats, keep, packed, offsets, _ = transform.encode_on_device(codec, eb, cem, x)
torch.cuda.synchronize(); codec.synchronize()
def ev_time(fn, stream=None):
    s = stream or torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        a.record(s); fn(); b.record(s)
    torch.cuda.synchronize()
    return a.elapsed_time(b)
for i in range(3):
    print("encode_dev alone on main: %.3f ms" % ev_time(lambda: cem.encode_dev(iv)))
side = codec.coder_stream()
for i in range(2):
    print("encode_dev alone on side: %.3f ms" % ev_time(lambda: cem.encode_dev(iv), side))
# right after the transforms of 64 cubes on the same stream
for i in range(2):
    codec.analysis(x[:64])
    print("encode_dev after analysis(64), same stream: %.3f ms" % ev_time(lambda: cem.encode_dev(iv)))
# fresh intervals written just before
for i in range(2):
    iv2, *_ = transform.encode_on_device(codec, eb, cem, x)[:1]
    torch.cuda.synchronize()
    print("encode_dev of fresh intervals (after sync): %.3f ms" % ev_time(lambda: cem.encode_dev(iv2)))
    print("   same intervals again: %.3f ms" % ev_time(lambda: cem.encode_dev(iv2)))
    print("   original intervals: %.3f ms" % ev_time(lambda: cem.encode_dev(iv)))
print("iv equal:", bool((iv2 == iv).all()))

# ---- inside the real compress path
from pcgcv1_b200.models import model_voxception
pinned = torch.from_numpy(cubes).pin_memory()
for variant in ("host cubes", "device cubes"):
    src = pinned if variant == "host cubes" else x
    for it in range(3):
        codec.profile(True); codec.profile_report()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = transform._compress_hyper_gpu_coder(codec, eb, cem, src, False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        rep = codec.profile_report(); codec.profile(False)
        enc = [r for r in rep if r["tag"].startswith("range_")]
        print("%s: compress wall %.1f ms; %s" % (variant, dt, [(r["tag"], r["count"], round(r["ms"], 3)) for r in enc]))
