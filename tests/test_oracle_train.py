"""The training oracle (oracle/train.py) checks itself on the CPU: its autograd gradients agree with central finite differences of
its own loss, and its Adam restatement reproduces the closed form of the first step."""
import numpy as np

from oracle import train as otrain
from pcgcv1_b200 import synthetic, weights as W


def test_oracle_gradient_matches_finite_differences_and_adam_first_step():
    w = W.synthetic_weights("voxception")
    cubes, _ = synthetic.surface_cubes(1, seed=4)
    import torch
    kw = dict(seed=3, entropy_dtype=torch.float64)          # smooth loss for the finite differences
    terms, grads, _ = otrain.forward_backward(w, cubes, **kw)
    assert np.isfinite(terms["loss"]) and terms["bpp_ae"] > 0 and terms["bpp_hyper"] > 0
    key = "synthesis_transform/deconv_out/bias"
    eps = 1e-4
    wp, wm = dict(w), dict(w)
    wp[key] = np.asarray(w[key], np.float64) + eps
    wm[key] = np.asarray(w[key], np.float64) - eps
    lp = otrain.forward_backward(wp, cubes, **kw)[0]["loss"]
    lm = otrain.forward_backward(wm, cubes, **kw)[0]["loss"]
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - grads[key][0]) <= 1e-5 * max(1.0, abs(fd))
    # Adam: after the first step every coordinate moves by lr * sign(g) (up to eps)
    g = grads[key]
    p, m, v = otrain.adam_update(np.zeros_like(g), g, np.zeros_like(g), np.zeros_like(g), 1, lr=1e-5)
    assert np.allclose(p, -1e-5 * np.sign(g), rtol=1e-6, atol=1e-12)
