import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import train as ot
from pcgcv1_b200 import runtime, synthetic, training, weights as W
codec = runtime.get_codec("voxception", "")
w = W.synthetic_weights("voxception")
cubes, _ = synthetic.surface_cubes(1, seed=4)
tr = training.HyperTrainer(codec, w)
ops = tr.ops
x = codec.to_device(cubes, torch.float32)
with torch.no_grad():
    y = tr._run("analysis_transform", x)
    z = tr._run("hyper_encoder", y)
stats = {}
zt = training._FactorizedRate.apply(z, tr.params["estimator/matrix"], tr.params["estimator/bais"], tr.params["estimator/factor"], -0.01, 3, 1e-9, ops, stats)
with torch.no_grad():
    loc, s_raw = tr._hyper_decoder(zt)
scale = torch.clamp(s_raw.abs(), min=1e-9)
print("scale min/max", float(scale.min()), float(scale.max()), "loc", float(loc.min()), float(loc.max()), "y", float(y.min()), float(y.max()))
yl = y.detach().clone().requires_grad_(True); ll = loc.detach().clone().requires_grad_(True); sl = scale.detach().clone().requires_grad_(True)
coef = -0.0005
yt = training._LaplaceRate.apply(yl, ll, sl, coef, 3, 1e-9, ops, stats)
yt.backward(torch.zeros_like(yt))
codec.synchronize()
# autograd fp64 on the same y_t, loc, scale
Y = torch.tensor(yt.detach().cpu().numpy().astype(np.float64), requires_grad=True)
L = torch.tensor(loc.cpu().numpy().astype(np.float64), requires_grad=True)
S = torch.tensor(scale.cpu().numpy().astype(np.float64), requires_grad=True)
p = torch.clamp(ot.laplace_likelihood(Y, L, S), min=1e-9)
(coef * torch.log(p).sum()).backward()
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for name, a, b in (("gy", yl.grad, Y.grad), ("gloc", ll.grad, L.grad), ("gscale", sl.grad, S.grad)):
    a = a.cpu().numpy().astype(np.float64); b = b.numpy()
    d = np.abs(a - b)
    i = np.unravel_index(np.argmax(d), d.shape)
    print(name, "rel", rel(a, b), "worst at", i, "gpu", a[i], "ref", b[i], "y_t", Y.detach().numpy()[i], "loc", L.detach().numpy()[i], "scale", S.detach().numpy()[i], "p", float(p.detach().numpy()[i]))
    print("   frac of elements with |d| > 1e-3*max:", float((d > 1e-3 * np.abs(b).max()).mean()))
print("logsum check", float(stats["bits_y"].sum()) * -math.log(2), float(torch.log(p).sum()))
