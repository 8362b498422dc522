"""Dev tool: the GPU range coder kernels timed alone and beside a stream of conv kernels (co-residency check)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgcv1_b200 import runtime, synthetic, transform
from pcgcv1_b200.models.conditional_entropy_model import SymmetricConditional

B = int(sys.argv[1]) if len(sys.argv) > 1 else 191
codec = runtime.get_codec("voxception", "")
cubes, _, nums = synthetic.workload("vox10", seed=0, max_cubes=B)
B = len(cubes)
x = codec.to_device(cubes)
eb = transform._bottleneck(codec, 8)
cem = SymmetricConditional().bind(codec)
iv, mm_all, z_hats, keep, packed, offsets, _ = transform.encode_on_device(codec, eb, cem, x, keep_side_info=True)
torch.cuda.synchronize()
codec.synchronize()
mm = mm_all.cpu().numpy()
locs = torch.cat([k[0] for k in keep]).reshape(B, -1); scales = torch.cat([k[1] for k in keep]).reshape(B, -1)
print("B", B, "bytes/cube", float(offsets[-1]) / B, "N range", (mm[:, 1] - mm[:, 0] + 1).min(), (mm[:, 1] - mm[:, 0] + 1).max())

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

print("encode alone       %.3f ms" % timeit(lambda: cem.encode_dev(iv)))
off = codec.row_offsets(mm, 65536)
hdr = codec.to_device(np.concatenate([off, mm.reshape(-1).astype(np.int64)]))
off_d, mm_d = hdr[:B + 1], hdr[B + 1:].to(torch.int32)
rows = codec.laplace_cdf_dev(locs, scales, mm_d, off_d, int(off[-1]))
nmax = int((mm[:, 1] - mm[:, 0] + 1).max())
print("cdf rows alone     %.3f ms" % timeit(lambda: codec.laplace_cdf_dev(locs, scales, mm_d, off_d, int(off[-1]))))
print("decode alone       %.3f ms" % timeit(lambda: codec.gpu_range_decode(packed, offsets, rows, off_d, int(off[-1]), mm_d, nmax, B, 65536)))
for nb in (1, 16, 48):
    print("decode alone B=%-3d  %.3f ms" % (nb, timeit(lambda: codec.gpu_range_decode(packed, offsets[:nb + 1], rows, off_d[:nb + 1], int(off[nb]), mm_d[:2 * nb], nmax, nb, 65536))))
    print("encode alone B=%-3d  %.3f ms" % (nb, timeit(lambda: cem.encode_dev(iv[:nb]))))

# beside conv kernels: analysis of 64 cubes in a loop on the main stream, coder on the side stream
side = codec.coder_stream()
y64 = x[:64]
def conv_loop(n): 
    for _ in range(n): codec.analysis(y64)
t_conv = timeit(lambda: conv_loop(1))
print("analysis(64) alone %.3f ms" % t_conv)
def both(fn_side):
    ev0 = torch.cuda.Event(); ev0.record()
    with torch.cuda.stream(side):
        side.wait_event(ev0)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(side); fn_side(); s1.record(side)
    conv_loop(2)
    torch.cuda.current_stream().wait_stream(side)
    return s0, s1
for name, fn in (("encode", lambda: cem.encode_dev(iv)), ("decode", lambda: codec.gpu_range_decode(packed, offsets, rows, off_d, int(off[-1]), mm_d, nmax, B, 65536))):
    both(fn); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s0, s1 = both(fn); b.record(); torch.cuda.synchronize()
    print("%s beside 2x analysis(64): total %.3f ms (2x analysis alone %.3f), coder kernel %.3f ms" % (name, a.elapsed_time(b), 2 * t_conv, s0.elapsed_time(s1)))
