"""Oracle of one training step (test infrastructure only): the reference's graph ``train_hyper.py:184-214`` + ``loss.py:8-33``
(+ ``get_focal_loss`` ``loss.py:83-93``, ``get_classify_metrics`` ``:60-77``; the three loss functions are pinned by
tests/golden/golden_loss.npz = the reference's own loss.py executed by tests/golden/make_golden_loss.py)
restated with torch-CPU autograd in float64 (or float32), the gradients coming from ``loss.backward()`` instead of
``tf.GradientTape``.  Noise = the Philox stream of oracle/entropy.py (the CUDA path draws the same numbers; the reference itself
is unseeded).  Adam restates ``tf.train.AdamOptimizer`` (TF 1.13 ``training/adam.py``: lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
var -= lr_t * m / (sqrt(v) + eps))."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import entropy, nets

LN2 = math.log(2.0)


def _net(P: Dict[str, torch.Tensor], net: str) -> Dict[str, torch.Tensor]:
    pre = net + "/"
    return {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}


def bottleneck_likelihood(x_cl: torch.Tensor, P: Dict[str, torch.Tensor]) -> torch.Tensor:
    """EntropyBottleneck._likelihood (entropy_model.py:72-151) on a channels-last tensor, torch autograd."""
    C = x_cl.shape[-1]
    x = x_cl.reshape(-1, C).t().reshape(C, 1, -1)

    def logits(v):
        for i in range(4):
            m = torch.nn.functional.softplus(P["estimator/matrix_%d" % i])
            v = torch.matmul(m, v) + P["estimator/bais_%d" % i]
            v = v + torch.tanh(P["estimator/factor_%d" % i]) * torch.tanh(v)
        return v
    lower, upper = logits(x - 0.5), logits(x + 0.5)
    sign = -torch.sign(lower + upper).detach()
    p = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    return p.reshape(C, -1).t().reshape(x_cl.shape)


def laplace_likelihood(x, loc, scale):
    """SymmetricConditional._likelihood (conditional_entropy_model.py:21-56), torch autograd (tf.sign has no gradient)."""
    def cum(t):
        e = torch.exp(-torch.abs(t - loc) / scale)
        return torch.where(t <= loc, 0.5 * e, 1.0 - 0.5 * e)
    upper, lower = x + 0.5, x - 0.5
    sign = torch.sign(upper + lower - loc).detach()
    upper = -sign * (upper - loc) + loc
    lower = -sign * (lower - loc) + loc
    return torch.abs(cum(upper) - cum(lower))


def bce_loss(pred, label):
    """get_bce_loss (loss.py:8-33)."""
    occ = torch.clamp(torch.sigmoid(pred), 1e-7, 1.0 - 1e-7)
    neg, pos = label[..., 0] == 0, label[..., 0] > 0
    o = occ[..., 0]
    return (-torch.log(1.0 - o[neg])).mean(), (-torch.log(o[pos])).mean()


def focal_loss(prob, label, gamma=2.0, alpha=0.9):
    """get_focal_loss (loss.py:83-93) on probabilities: pt_1 = clip(where(label == 1, p, 1), 1e-3, .999), pt_0 = clip(where(label
    == 0, p, 0), 1e-3, .999); -sum(alpha (1 - pt_1)^gamma log pt_1) - sum((1 - alpha) pt_0^gamma log(1 - pt_0)).  torch.clamp has
    K.clip's gradient (zero outside the bounds); the constant branches of the two where() calls stay in the value."""
    pt_1 = torch.clamp(torch.where(label == 1, prob, torch.ones_like(prob)), 1e-3, 0.999)
    pt_0 = torch.clamp(torch.where(label == 0, prob, torch.zeros_like(prob)), 1e-3, 0.999)
    return -(alpha * (1.0 - pt_1) ** gamma * torch.log(pt_1)).sum(), -((1.0 - alpha) * pt_0 ** gamma * torch.log(1.0 - pt_0)).sum()


def classify_metrics(pred, label, th=0.0):
    """get_classify_metrics (loss.py:60-77): precision, recall, IoU of (pred > th) against (label > th)."""
    p, l = (pred > th).double(), (label > th).double()
    tp, fp, fn = (p * l).sum(), (p * (1 - l)).sum(), ((1 - p) * l).sum()
    return float(tp / (tp + fp)), float(tp / (tp + fn)), float(tp / (tp + fp + fn))


def forward_backward(weights: Dict[str, np.ndarray], cubes: np.ndarray, seed: int = 0, alpha=0.75, beta=3.0, gamma=1.0, delta=1.0,
                     lower_bound=1e-9, likelihood_bound=1e-9, dtype=torch.float64, noise_dtype=np.float32, entropy_dtype=torch.float32,
                     distortion="bce", focal_gamma=2.0, focal_alpha=0.9):
    """-> (terms dict of floats, grads dict name -> np.ndarray keyed like the weight file).

    ``dtype`` is the arithmetic of the transforms; ``entropy_dtype`` that of the two likelihood formulas.  The reference runs
    everything in float32, and its conditional likelihood is NOT precision-neutral: ``sign(upper + lower - loc)`` is
    ``sign(2x - loc)`` (conditional_entropy_model.py:47), so for x right of loc with 2x < loc nothing is reflected, both CDF
    values round to 1.0f and the likelihood collapses to the 1e-9 floor with a ZERO gradient -- where float64 sees a tiny
    likelihood with a huge gradient.  The oracle therefore evaluates the likelihoods in float32 (what TF does) and keeps
    float64 for the convolutions and the reductions."""
    P = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in weights.items()
         if k.split("/")[0] in ("analysis_transform", "synthesis_transform", "hyper_encoder", "hyper_decoder", "estimator")}
    x = torch.tensor(cubes.astype(np.float64), dtype=dtype)
    y = nets.analysis_voxception(x, _net(P, "analysis_transform"))
    z = nets.hyper_encoder(y, _net(P, "hyper_encoder"))
    nz = entropy.philox_uniform(seed, z.numel()).astype(noise_dtype).reshape(z.shape)
    z_t = z + torch.tensor(nz, dtype=dtype)
    ed = entropy_dtype
    Pe = {k: v.to(ed) for k, v in P.items() if k.startswith("estimator/")}
    p_z = torch.clamp(bottleneck_likelihood(z_t.to(ed), Pe), min=likelihood_bound).to(dtype)
    loc, scale = nets.hyper_decoder(z_t, _net(P, "hyper_decoder"))
    scale = torch.clamp(scale, min=lower_bound)
    ny = entropy.philox_uniform((seed + entropy.Y_NOISE_SEED_OFFSET) & 0xFFFFFFFFFFFFFFFF, y.numel()).astype(noise_dtype).reshape(y.shape)
    y_t = y + torch.tensor(ny, dtype=dtype)
    p_y = torch.clamp(laplace_likelihood(y_t.to(ed), loc.to(ed), scale.to(ed)), min=likelihood_bound).to(dtype)
    x_t = nets.synthesis_voxception(y_t, _net(P, "synthesis_transform"))
    num_points = float((x.sum(-1) > 0).sum())
    bpp_ae = torch.log(p_y).sum() / (-LN2 * num_points)
    bpp_hyper = torch.log(p_z).sum() / (-LN2 * num_points)
    if distortion == "focal":           # loss.py:83-93 on sigmoid(x_tilde); the reference defines it, BASELINE config 5 names it
        f_full, f_empty = focal_loss(torch.sigmoid(x_t), x, focal_gamma, focal_alpha)
        dist = f_full + f_empty
        parts = (("focal_full", f_full), ("focal_empty", f_empty))
    else:
        zeros, ones = bce_loss(x_t, x)
        dist = beta * zeros + 1.0 * ones
        parts = (("zeros", zeros), ("ones", ones))
    loss = alpha * dist + delta * bpp_ae + gamma * bpp_hyper
    loss.backward()
    terms = {k: float(v.detach()) for k, v in parts + (("distortion", dist), ("bpp_ae", bpp_ae), ("bpp_hyper", bpp_hyper), ("loss", loss))}
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros(v.shape)) for k, v in P.items()}
    return terms, grads, {"y": y.detach().numpy(), "x_tilde": x_t.detach().numpy()}


def adam_update(p, g, m, v, t, lr=1e-5, beta1=0.9, beta2=0.999, eps=1e-8):
    """One tf.train.AdamOptimizer update in float64; returns (p, m, v)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v
