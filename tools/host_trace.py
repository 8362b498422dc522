"""Dev tool: HOST-side timeline of one end-to-end step (which Python call blocks for how long), by wrapping the pipeline's
building blocks with wall-clock timers.  Complements tools/timeline.py (the GPU side)."""
import os, sys, time, functools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import model_voxception
from pcgcv1_b200.models.conditional_entropy_model import SymmetricConditional
from pcgcv1_b200.models.entropy_model import EntropyBottleneck

T0 = [0.0]
LOG = []
ON = [False]


def wrap(owner, name, label=None):
    fn = getattr(owner, name)
    lab = label or name

    @functools.wraps(fn)
    def w(*a, **k):
        if not ON[0]:
            return fn(*a, **k)
        t = time.perf_counter()
        r = fn(*a, **k)
        LOG.append(((t - T0[0]) * 1e3, (time.perf_counter() - t) * 1e3, lab))
        return r
    setattr(owner, name, w)


wrap(transform, "encode_on_device")
wrap(EntropyBottleneck, "compress_begin")
wrap(EntropyBottleneck, "compress_finish")
wrap(EntropyBottleneck, "decompress_progressive")
wrap(runtime, "to_host")
wrap(runtime, "host_copy")
wrap(runtime.Codec, "upload_strings")
wrap(runtime.Codec, "synchronize")
wrap(runtime.Codec, "hyper_decode")
wrap(runtime.Codec, "synthesis")
wrap(runtime.Codec, "analysis")
wrap(runtime.Codec, "topk")
wrap(runtime.Codec, "to_device")
wrap(runtime.ProgressiveDecode, "wait", "z wait")
wrap(SymmetricConditional, "decode_dev")
wrap(SymmetricConditional, "encode_dev")
wrap(SymmetricConditional, "intervals_dev")
wrap(transform, "_as_list_of_bytes")

cubes, _, nums = synthetic.workload("vox10", seed=0)
codec = runtime.get_codec("voxception", "")
pinned = torch.from_numpy(cubes).pin_memory()


def dump(name, wall):
    print("%s: %.1f ms" % (name, wall))
    for t, lab in transform.TRACE:
        print("   %7.2f           mark: %s" % ((t - T0[0]) * 1e3, lab))
    transform.TRACE.clear()
    for t, d, lab in LOG:
        if d >= 0.05:
            print("   %7.2f  +%6.2f  %s" % (t, d, lab))
    LOG.clear()


transform.TRACE = []
for it in range(4):
    ON[0] = it == 3
    transform.TRACE.clear()
    torch.cuda.synchronize(); T0[0] = t0 = time.perf_counter()
    out = transform.compress_hyper(pinned, model_voxception, "")
    ta = time.perf_counter()
    host = [o.numpy() for o in out]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if ON[0]:
        LOG.append(((ta - t0) * 1e3, (t1 - ta) * 1e3, ".numpy() of the stream fields"))
        T0[0] = t0
        dump("compress", (t1 - t0) * 1e3)
    T0[0] = t1
    xs = transform.decompress_hyper(*host, model_voxception, "")
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if ON[0]:
        dump("decompress", (t2 - t1) * 1e3)
    T0[0] = t2
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if ON[0]:
        dump("select", (t3 - t2) * 1e3)
