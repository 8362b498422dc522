"""Drop-in for the reference's ``transform.py``: ``compress_factorized`` / ``decompress_factorized`` /
``compress_hyper`` / ``decompress_hyper`` with the reference's signatures and return tuples
(transform.py:24-56,58-87,91-197,200-259).  Every returned value has ``.numpy()`` because the
callers (test.py:81,89,101-103,115; eval.py:56-57,87-89) call it.

What changes underneath: the per-cube ``tf.map_fn(..., parallel_iterations=1)`` loops
(transform.py:48,84,122,131,143,166,183,193,230,247,256) become one batched call per net into
libpcgc_b200.so; quantisation / likelihood / per-element CDF rows are fused GPU kernels; only the
range coder runs on the host (thread pool over cubes).  The hyper decoder is bit-reproducible, so a
stream written by ``compress_hyper`` decodes with ``decompress_hyper`` (the reference's GPU path
does not guarantee that: README.md:111-114).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import runtime
from .models.conditional_entropy_model import SymmetricConditional
from .models.entropy_model import EntropyBottleneck

_VERBOSE = bool(int(os.environ.get("PCGC_VERBOSE", "0")))


def _log(msg, start=None):
    if _VERBOSE:
        if start is not None:
            torch.cuda.synchronize()
            print("{}: {}s".format(msg, round(time.time() - start, 4)))
        else:
            print(msg)


def _strings_array(strings):
    a = np.empty(len(strings), dtype=object)
    for i, s in enumerate(strings):
        a[i] = s
    return a


def _as_list_of_bytes(strings):
    strings = runtime.unwrap(strings)
    if isinstance(strings, (bytes, bytearray)):
        return [bytes(strings)]
    return [bytes(s) for s in list(strings)]


def _bottleneck(codec, channels):
    return EntropyBottleneck().bind(codec, codec.bottleneck_slot(channels))


# ---------------------------------------------------------------- factorized entropy model
def compress_factorized(cubes, model, ckpt_dir):
    """cubes [B,64,64,64,1] -> (strings, min_v, max_v, shape)  (transform.py:24-56)."""
    _log("===== Compress =====")
    codec = runtime.get_codec(model, ckpt_dir)
    x = codec.to_device(cubes)
    start = time.time()
    ys = codec.analysis(x)
    _log("Analysis Transform", start)
    start = time.time()
    strings, min_v, max_v = _bottleneck(codec, ys.shape[-1]).compress(ys)
    shape = runtime.HostResult(np.array(ys.shape, dtype=np.int32))
    _log("Entropy Encode", start)
    return strings, min_v, max_v, shape


def decompress_factorized(strings, min_v, max_v, shape, model, ckpt_dir):
    """-> xs [B,64,64,64,1] occupancy logits  (transform.py:58-87)."""
    _log("===== Decompress =====")
    codec = runtime.get_codec(model, ckpt_dir)
    shape = np.asarray(runtime.unwrap(shape)).reshape(-1)
    start = time.time()
    ys = _bottleneck(codec, int(shape[-1])).decompress(strings, min_v, max_v, shape, shape[-1])
    _log("Entropy Decode", start)
    start = time.time()
    xs = codec.synthesis(ys.tensor)
    _log("Synthesis Transform", start)
    return runtime.DeviceResult(xs)


# ---------------------------------------------------------------- hyperprior (conditional) model
def compress_hyper(cubes, model, ckpt_dir, decompress=False):
    """cubes [B,64,64,64,1] -> (y_strings[B], y_min_vs[B], y_max_vs[B], y_shape, z_strings, z_min_v,
    z_max_v, z_shape[, x_decodeds])  (transform.py:91-197)."""
    _log("===== Compress =====")
    codec = runtime.get_codec(model, ckpt_dir)
    entropy_bottleneck = _bottleneck(codec, 8)
    conditional_entropy_model = SymmetricConditional().bind(codec)
    x = codec.to_device(cubes)

    start = time.time()
    ys = codec.analysis(x)
    _log("Analysis Transform", start)
    start = time.time()
    zs = codec.hyper_encode(ys)
    _log("Hyper Encoder", start)
    z_hats, _, _, _ = codec.factorized(entropy_bottleneck._slot, zs, want_p=False, want_bits=False)
    start = time.time()
    locs, scales = codec.hyper_decode(z_hats, 1e-9)                      # lower_bound = 1e-9, transform.py:145-146
    _log("Hyper Decoder", start)
    start = time.time()
    z_strings, z_min_v, z_max_v = entropy_bottleneck.compress(zs)
    z_shape = runtime.HostResult(np.array(zs.shape, dtype=np.int32))
    _log("Entropy Encode (Hyper)", start)
    start = time.time()
    strings, y_min_vs, y_max_vs = conditional_entropy_model.compress_cubes(ys, locs, scales)
    y_shape = runtime.HostResult(np.array((1,) + tuple(ys.shape[1:]), dtype=np.int64))
    _log("Entropy Encode", start)
    out = (runtime.HostResult(_strings_array(strings)), runtime.HostResult(y_min_vs.astype(np.int32)),
           runtime.HostResult(y_max_vs.astype(np.int32)), y_shape, z_strings, z_min_v, z_max_v, z_shape)
    if decompress:
        start = time.time()
        y_dec = conditional_entropy_model.decompress_cubes(strings, locs, scales, y_min_vs, y_max_vs)
        _log("Entropy Decode", start)
        start = time.time()
        x_dec = codec.synthesis(y_dec.reshape(ys.shape))
        _log("Synthesis Transform", start)
        return out + (runtime.DeviceResult(x_dec),)
    return out


def decompress_hyper(y_strings, y_min_vs, y_max_vs, y_shape, z_strings, z_min_v, z_max_v, z_shape, model, ckpt_dir):
    """-> xs [B,64,64,64,1] occupancy logits  (transform.py:200-259)."""
    _log("===== Decompress =====")
    codec = runtime.get_codec(model, ckpt_dir)
    entropy_bottleneck = _bottleneck(codec, 8)
    conditional_entropy_model = SymmetricConditional().bind(codec)
    z_shape = np.asarray(runtime.unwrap(z_shape)).reshape(-1)
    y_shape = [int(v) for v in np.asarray(runtime.unwrap(y_shape)).reshape(-1)]

    start = time.time()
    zs = entropy_bottleneck.decompress(z_strings, z_min_v, z_max_v, z_shape, z_shape[-1])
    _log("Entropy Decoder (Hyper)", start)
    start = time.time()
    locs, scales = codec.hyper_decode(zs.tensor, 1e-9)
    _log("Hyper Decoder", start)
    start = time.time()
    strings = _as_list_of_bytes(y_strings)
    B = locs.shape[0]
    if len(strings) != B:
        raise ValueError("got %d y strings for %d cubes" % (len(strings), B))
    ys = conditional_entropy_model.decompress_cubes(strings, locs, scales, np.asarray(runtime.unwrap(y_min_vs)),
                                                    np.asarray(runtime.unwrap(y_max_vs)))
    _log("Entropy Decoder", start)
    start = time.time()
    xs = codec.synthesis(ys.reshape([B] + y_shape[1:]))
    _log("Synthesis Transform", start)
    return runtime.DeviceResult(xs)
