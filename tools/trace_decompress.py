"""Dev tool: host-side timeline of decompress_hyper's chunk pipeline (where the host waits, when GPU work is issued)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.models import model_voxception
from pcgcv1_b200.models.conditional_entropy_model import SymmetricConditional

cubes, pos, nums = synthetic.workload("vox10")
pinned = torch.from_numpy(cubes).pin_memory()
codec = runtime.get_codec("voxception", "")
for _ in range(2):
    out = transform.compress_hyper(pinned, model_voxception, "")
    host = [o.numpy() for o in out]
    transform.decompress_hyper(*host, model_voxception, "")
torch.cuda.synchronize()
y_strings, y_min_vs, y_max_vs, y_shape, z_strings, z_min_v, z_max_v, z_shape = host
eb = transform._bottleneck(codec, 8)
cem = SymmetricConditional().bind(codec)
T0 = time.perf_counter()
def t(msg):
    print("%7.2f ms  %s" % ((time.perf_counter() - T0) * 1e3, msg))
z_get = eb.decompress_progressive(z_strings, z_min_v, z_max_v, z_shape, 8)
t("z decode started")
strings = [bytes(s) for s in y_strings]
B = len(strings)
chunks = transform._chunks(B, small_first=True)
pending = None
parts = []
ev_done = []
for k, (a, b) in enumerate(chunks):
    zc = z_get(a, b); t("chunk %d [%d,%d): z ready" % (k, a, b))
    locs, scales = codec.hyper_decode(zc, 1e-9)
    stage, done, off, mm, E = cem.decode_begin(locs, scales, y_min_vs[a:b], y_max_vs[a:b], k % 2); t("chunk %d: HD + rows kernel done (host synced), D2H queued" % k)
    job = transform._pool().submit(cem.decode_finish, strings[a:b], stage, done, off, mm, E, k % 2)
    if pending is not None:
        (pa, pb), pj = pending
        yh = pj.result(); t("chunk %d: host decode finished" % (k - 1))
        ys = codec.to_device(yh).reshape([pb - pa, 16, 16, 16, 16])
        parts.append(codec.synthesis(ys)); t("chunk %d: synthesis queued" % (k - 1))
        e = torch.cuda.Event(); e.record(); ev_done.append(e)
    pending = ((a, b), job)
(pa, pb), pj = pending
yh = pj.result(); t("last chunk: host decode finished")
ys = codec.to_device(yh).reshape([pb - pa, 16, 16, 16, 16])
parts.append(codec.synthesis(ys)); t("last chunk: synthesis queued")
torch.cuda.synchronize(); t("GPU idle (all done)")
