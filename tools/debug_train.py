import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import train as otrain
from pcgcv1_b200 import runtime, synthetic, training, weights as W
codec = runtime.get_codec("voxception", "")
w = W.synthetic_weights("voxception")
cubes, _ = synthetic.surface_cubes(1, seed=4)
tr = training.HyperTrainer(codec, w)
out = tr.forward_backward(cubes, seed=3)
print(tr.loss_terms(out))
ref_terms, ref_grads, _ = otrain.forward_backward(w, cubes, seed=3)
print(ref_terms)
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for k, p in tr.params.items():
    if k.startswith("estimator/"):
        name = k.split("/")[1]
        ref = np.concatenate([ref_grads["estimator/%s_%d" % (name, i)].reshape(-1) for i in range(4)])
    else:
        ref = ref_grads[k]
    g = p.grad.detach().cpu().numpy()
    e = rel(g, ref)
    if e > 1e-4 or "kernel" in k and ("conv_in" in k or "out" in k or "deconv4" in k or "conv3" in k or "up_" in k or "down_" in k):
        print("%-55s rel %.3e  |ref|max %.3e" % (k, e, np.abs(ref).max()))
