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


def test_oracle_loss_functions_match_the_reference_loss_py():
    """golden_loss.npz = /root/reference/loss.py executed by tests/golden/make_golden_loss.py (BCE means, focal sums, metrics)."""
    import os
    import torch
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_loss.npz"))
    for tag in ("a", "b"):
        # float32 like the reference: case "b" holds saturated logits, where the float32 clip bound 1 - 1e-7 -> 1 - 2^-23 decides the value
        pred, label = torch.tensor(g[tag + "_pred"], dtype=torch.float32), torch.tensor(g[tag + "_label"], dtype=torch.float32)
        empty, full = otrain.bce_loss(pred, label)
        assert np.allclose([float(empty), float(full)], g[tag + "_bce"], rtol=2e-6)
        # the reference function takes PROBABILITIES: float32 sigmoid as the golden script fed it
        prob = torch.tensor((np.float32(1) / (np.float32(1) + np.exp(-g[tag + "_pred"]))).astype(np.float32), dtype=torch.float64)
        label = label.double()
        for key, kw in (("_focal", {}), ("_focal_g3_a75", dict(gamma=3.0, alpha=0.75))):
            f1, f0 = otrain.focal_loss(prob, label, **kw)
            assert abs(float(f1 + f0) - g[tag + key][0]) <= 3e-6 * g[tag + key][0], (tag, key)
        assert np.allclose(otrain.classify_metrics(pred, label), g[tag + "_metrics"], rtol=1e-6)


def test_oracle_focal_training_loss_has_consistent_gradient():
    import torch
    w = W.synthetic_weights("voxception")
    cubes, _ = synthetic.surface_cubes(1, seed=4)
    kw = dict(seed=3, entropy_dtype=torch.float64, distortion="focal")
    terms, grads, _ = otrain.forward_backward(w, cubes, **kw)
    assert abs(terms["distortion"] - (terms["focal_full"] + terms["focal_empty"])) < 1e-9 * terms["distortion"]
    key = "synthesis_transform/deconv_out/bias"
    eps = 1e-4
    wp, wm = dict(w), dict(w)
    wp[key] = np.asarray(w[key], np.float64) + eps
    wm[key] = np.asarray(w[key], np.float64) - eps
    fd = (otrain.forward_backward(wp, cubes, **kw)[0]["loss"] - otrain.forward_backward(wm, cubes, **kw)[0]["loss"]) / (2 * eps)
    assert abs(fd - grads[key][0]) <= 1e-5 * max(1.0, abs(fd))
