"""Weight containers for the drop-in transforms (host side, NumPy only).

The reference restores ``tf.train.Checkpoint(analysis_transform=..., synthesis_transform=...,
hyper_encoder=..., hyper_decoder=..., estimator=...)`` (``transform.py:107-112``); with
``ckpt_dir == ''`` nothing is restored and the Keras default init stays.  Here a checkpoint
directory holds one ``weights.npz`` whose keys are ``<top-level key>/<keras layer name>/kernel``
(or ``/bias``) in the Keras layouts, plus ``estimator/matrix_i|bais_i|factor_i``
(``models/entropy_model.py:51-66``, the misspelt ``bais`` is the reference's variable name).

``ckpt_dir == ''`` gives SEEDED synthetic weights (documented deviation: the reference is
unseeded, and its default glorot init makes every latent round to 0 -- SURVEY.md section 7.0 --
which ``pmf_to_quantized_cdf`` rejects).  The synthetic init is He-normal on the hidden layers
with the last layer of every net rescaled by constants calibrated once on the synthetic
vox10 cloud (``tools/calibrate_weights.py``) so that y, z stay inside the bitstream format's
[-15, 15] symbol range with a useful spread.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

from . import netspec

Weights = Dict[str, np.ndarray]

# (kernel gain, bias std) of the final layer(s) of each net; calibrated by tools/calibrate_weights.py
CALIBRATION = {
    "voxception": {
        "analysis_transform/conv_out": (0.55, 0.1),
        "hyper_encoder/conv3": (1.5, 0.1),
        "hyper_decoder/deconv4_1": (0.5, 0.1),
        "hyper_decoder/deconv4_2": (0.3, 0.05),
        "synthesis_transform/deconv_out": (0.25, 0.1),
    },
    "simple": {
        "analysis_transform/conv_3": (5.5, 0.0),
        "synthesis_transform/deconv_3": (0.6, 0.1),
    },
}


def entropy_bottleneck_params(channels: int, rng: np.random.Generator, init_scale: float = 8.0,
                              filters=(3, 3, 3), trained_like: bool = True) -> Weights:
    """EntropyBottleneck.build (entropy_model.py:25-70).  ``trained_like`` perturbs matrices and
    factors away from their constant/zero init so every term of the density is exercised."""
    f = (1,) + tuple(filters) + (1,)
    scale = init_scale ** (1.0 / (len(filters) + 1))
    p: Weights = {}
    for i in range(len(filters) + 1):
        init = np.log(np.expm1(1.0 / scale / f[i + 1]))
        m = np.full((channels, f[i + 1], f[i]), init, np.float32)
        b = rng.uniform(-0.5, 0.5, (channels, f[i + 1], 1)).astype(np.float32)
        fa = np.zeros((channels, f[i + 1], 1), np.float32)
        if trained_like:
            m = (m + rng.normal(0, 0.3, m.shape)).astype(np.float32)
            fa = rng.normal(0, 0.5, fa.shape).astype(np.float32)
        p["matrix_%d" % i] = m
        p["bais_%d" % i] = b
        p["factor_%d" % i] = fa
    return p


def synthetic_weights(model: str = "voxception", seed: int = 1234, calibration=None) -> Weights:
    """Flat dict ``<net>/<layer>/kernel|bias`` + ``estimator/*`` for ``model``."""
    cal = CALIBRATION[model] if calibration is None else calibration
    rng = np.random.default_rng(seed)
    w: Weights = {}
    for (m, net), layers in netspec.NETS.items():
        if m != model:
            continue
        for l in layers:
            shape = netspec.kernel_shape(l)
            fan_in = l.k ** 3 * l.cin
            if l.transposed:
                fan_in = fan_in / (l.stride ** 3)       # each output sees k^3/s^3 taps
            std = np.sqrt(2.0 / fan_in)
            bstd = 0.05
            last = l.name.endswith(("_conv1_2", "_conv2_3"))
            if last:
                std *= 0.5                              # keep the residual sum from blowing up
            key = "%s/%s" % (net, l.name)
            if key in cal:
                g, bstd = cal[key]
                std *= g
            w[key + "/kernel"] = rng.normal(0, std, shape).astype(np.float32)
            if l.bias:
                w[key + "/bias"] = rng.normal(0, bstd, (l.cout,)).astype(np.float32)
    ch = netspec.HYPER_CHANNELS if model == "voxception" else netspec.LATENT_CHANNELS[model]
    for k, v in entropy_bottleneck_params(ch, rng).items():
        w["estimator/" + k] = v
    if model == "voxception":
        # factorized mode with the voxception transforms uses a 16-channel bottleneck
        for k, v in entropy_bottleneck_params(netspec.LATENT_CHANNELS[model], rng).items():
            w["estimator_y/" + k] = v
    return w


def net_weights(w: Weights, net: str) -> Weights:
    """Sub-dict of one top-level checkpoint key with the prefix stripped."""
    p = net + "/"
    return {k[len(p):]: v for k, v in w.items() if k.startswith(p)}


def save(ckpt_dir: str, w: Weights) -> str:
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, "weights.npz")
    np.savez(path, **w)
    return path


def load(ckpt_dir: str, model: str = "voxception") -> Weights:
    """``ckpt_dir == ''`` -> seeded synthetic weights; else ``<ckpt_dir>/weights.npz``."""
    if not ckpt_dir:
        return synthetic_weights(model)
    path = os.path.join(ckpt_dir, "weights.npz")
    if not os.path.exists(path):
        from . import tf_checkpoint
        if tf_checkpoint.latest_checkpoint(ckpt_dir) is not None:         # a TF-1.13 checkpoint of the reference
            return tf_checkpoint.import_checkpoint(ckpt_dir, model)
        raise FileNotFoundError(
            "%s not found and no TF checkpoint (*.index) in that directory: this build reads weights.npz "
            "(keys <net>/<layer>/kernel) or the reference's TF-1.13 checkpoints (pcgcv1_b200/tf_checkpoint.py)" % path)
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
