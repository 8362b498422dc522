"""Dev tool: GPU timeline of one end-to-end step from the library's per-launch CUDA events (PCGC_PROF_TIMELINE form of
pcgc_profile_report): start and duration of every profiled launch group, the idle gaps of the device and the busy time per group."""
import os, sys, time, json, ctypes as C
os.environ["PCGC_PROF_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import model_voxception

cubes, _, nums = synthetic.workload("vox10", seed=0)
codec = runtime.get_codec("voxception", "")
pinned = torch.from_numpy(cubes).pin_memory()


def report():
    buf = C.create_string_buffer(1 << 20)
    codec._stream()
    codec._check(codec.lib.pcgc_profile_report(codec.ctx, buf, len(buf)))
    return json.loads(buf.value.decode())


def show(name, recs, wall_ms, verbose):
    if not recs:
        return
    iv = sorted((r["t0"], r["t0"] + r["ms"], r["tag"]) for r in recs)
    end = max(b for _, b, _ in iv)
    busy, cur_a, cur_b = 0.0, iv[0][0], iv[0][1]
    gaps = []
    for a, b, tag in iv[1:]:
        if a > cur_b:
            busy += cur_b - cur_a
            gaps.append((cur_b, a - cur_b, tag))
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    busy += cur_b - cur_a
    print("%s: wall %.1f ms, GPU span %.1f ms (first launch -> last end), union busy %.1f ms, %d launch groups" % (name, wall_ms, end, busy, len(iv)))
    for at, g, tag in sorted(gaps, key=lambda x: -x[1])[:12]:
        print("    idle %.2f ms at t=%.2f before %s" % (g, at, tag))
    if verbose:
        for a, b, tag in iv:
            print("    %8.3f %8.3f  %s" % (a, b - a, tag))


verbose = bool(int(os.environ.get("TL_VERBOSE", "0")))
for it in range(3):
    codec.profile(True); report()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = transform.compress_hyper(pinned, model_voxception, "")
    host = [o.numpy() for o in out]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rc = report()
    torch.cuda.synchronize(); t1b = time.perf_counter()
    xs = transform.decompress_hyper(*host, model_voxception, "")
    torch.cuda.synchronize(); t2 = time.perf_counter()
    rd = report()
    torch.cuda.synchronize(); t2b = time.perf_counter()
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
    torch.cuda.synchronize(); t3 = time.perf_counter()
    rs = report()
    codec.profile(False)
    if it == 2:
        show("compress", rc, (t1 - t0) * 1e3, verbose)
        show("decompress", rd, (t2 - t1b) * 1e3, verbose)
        show("select", rs, (t3 - t2b) * 1e3, verbose)
# without profiling
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = transform.compress_hyper(pinned, model_voxception, "")
    host = [o.numpy() for o in out]
    t1 = time.perf_counter()
    xs = transform.decompress_hyper(*host, model_voxception, "")
    t2 = time.perf_counter()
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
    torch.cuda.synchronize(); t3 = time.perf_counter()
print("unprofiled: compress %.1f ms, decompress %.1f ms, select %.1f ms, total %.1f ms -> %.0f cubes/s" %
      ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, len(cubes) / (t3 - t0)))
