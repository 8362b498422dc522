"""Golden vectors of the reference's loss functions, produced by EXECUTING /root/reference/loss.py in this container.

    python tests/golden/make_golden_loss.py     # needs /root/reference (read-only), writes golden_loss.npz here

``loss.py`` is imported unmodified; ``tensorflow`` / ``tensorflow.keras.backend`` are the NumPy float32 stand-ins below (TF 1.13
cannot be installed here).  Each stand-in is the documented element-wise meaning of the TF op, so what this pins is the
reference's FORMULAS: which voxels enter which mean, the clip bounds, the constant branches of the two ``tf.where`` calls of the
focal loss, sum against mean.  /root/reference does not exist on the GPU box: tests only read the .npz written here.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCGC_REFERENCE", "/root/reference")


def _where(cond, a=None, b=None):
    if a is None:
        return np.argwhere(cond)                       # tf.where(cond): int64 coordinates [n, rank]
    return np.where(cond, a, b)


def install():
    f32 = np.float32
    tf = types.ModuleType("tensorflow")
    tf.float32 = np.float32
    tf.enable_eager_execution = lambda: None
    tf.sigmoid = lambda x: (f32(1) / (f32(1) + np.exp(-np.asarray(x, f32)))).astype(f32)
    tf.clip_by_value = lambda x, lo, hi: np.clip(np.asarray(x, f32), f32(lo), f32(hi))
    tf.equal = lambda a, b: np.asarray(a) == b
    tf.greater = lambda a, b: np.asarray(a) > b
    tf.reduce_max = lambda x, axis=None: np.max(x, axis=axis)
    tf.reduce_mean = lambda x: np.mean(np.asarray(x, f32), dtype=np.float64).astype(f32)   # wide accumulator, float32 result
    tf.reduce_sum = lambda x: np.sum(np.asarray(x, f32), dtype=np.float64).astype(f32)
    tf.cast = lambda x, dtype: np.asarray(x).astype(np.dtype(dtype) if not isinstance(dtype, str) else dtype)
    tf.where = _where
    tf.gather_nd = lambda x, idx: np.asarray(x)[tuple(np.asarray(idx).T)]
    tf.negative = lambda x: -np.asarray(x)
    tf.log = lambda x: np.log(np.asarray(x, f32))
    tf.squeeze = lambda x, axis: np.squeeze(x, axis)
    tf.ones_like = np.ones_like
    tf.zeros_like = np.zeros_like
    keras = types.ModuleType("tensorflow.keras")
    K = types.ModuleType("tensorflow.keras.backend")
    K.clip = lambda x, lo, hi: np.clip(np.asarray(x, f32), f32(lo), f32(hi))
    K.pow = lambda x, a: np.power(np.asarray(x, f32), f32(a))
    K.log = lambda x: np.log(np.asarray(x, f32))
    K.sum = lambda x: np.sum(np.asarray(x, f32), dtype=np.float64).astype(f32)
    keras.backend = K
    tf.keras = keras
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.keras"] = keras
    sys.modules["tensorflow.keras.backend"] = K


def main():
    install()
    sys.path.insert(0, REF)
    ref = importlib.import_module("loss")
    rng = np.random.default_rng(20191005)
    out = {}
    for tag, shape, spread in (("a", (2, 8, 8, 8, 1), 3.0), ("b", (1, 6, 5, 4, 1), 12.0)):
        label = (rng.random(shape) < 0.2).astype(np.float32)
        pred = (rng.normal(0, spread, shape) + 4.0 * (label - 0.3)).astype(np.float32)
        if tag == "b":                                  # saturated logits on both sides of both clips (1e-7 / 1e-3)
            flat = pred.reshape(-1)
            flat[:8] = np.array([-40, 40, -17, 17, -7.5, 7.5, -6.9, 6.9], np.float32)
        empty, full = ref.get_bce_loss(pred, label)
        prob = (np.float32(1) / (np.float32(1) + np.exp(-pred))).astype(np.float32)
        out["%s_pred" % tag], out["%s_label" % tag] = pred, label
        out["%s_bce" % tag] = np.array([empty, full], np.float64)
        out["%s_focal" % tag] = np.array([ref.get_focal_loss(prob, label)], np.float64)                       # gamma=2, alpha=0.9
        out["%s_focal_g3_a75" % tag] = np.array([ref.get_focal_loss(prob, label, gamma=3, alpha=0.75)], np.float64)
        out["%s_metrics" % tag] = np.array(ref.get_classify_metrics(pred, label), np.float64)                # precision, recall, IoU
    np.savez_compressed(os.path.join(HERE, "golden_loss.npz"), **out)
    for k in sorted(out):
        if not k.endswith(("pred", "label")):
            print(k, out[k])


if __name__ == "__main__":
    main()
