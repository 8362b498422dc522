"""Drop-in for ``models/entropy_model.py``: ``EntropyBottleneck`` with the reference's constructor,
``__call__(inputs, training)``, ``compress(inputs)`` and ``decompress(strings, min_v, max_v, shape,
channels)`` (entropy_model.py:16-23,153-181,223-261,263-306).  Quantisation, the factorized density
and the pmf run in libpcgc_b200.so on the GPU; the 16-bit CDF normaliser and the range coder run in
its host part."""
from __future__ import annotations

import numpy as np
import torch

from .. import runtime


class EntropyBottleneck:
    def __init__(self, likelihood_bound=1e-9, range_coder_precision=16, init_scale=8, filters=(3, 3, 3),
                 codec=None, slot=None):
        if tuple(filters) != (3, 3, 3):
            raise NotImplementedError("the CUDA density kernel is specialised for filters=(3,3,3) (the reference's only setting)")
        self._likelihood_bound = float(likelihood_bound)
        self._range_coder_precision = int(range_coder_precision)
        self._init_scale = float(init_scale)
        self._codec = codec
        self._slot = slot

    def bind(self, codec, slot=None):
        self._codec, self._slot = codec, slot
        return self

    def _resolve(self, channels: int):
        if self._codec is None:
            self._codec = runtime.get_codec("voxception", "")
        if self._slot is None:
            self._slot = self._codec.bottleneck_slot(int(channels))
        return self._codec, self._slot

    def __call__(self, inputs, training=False, seed=0):
        """-> (quantised values, likelihoods), both shaped like ``inputs`` (entropy_model.py:153-181).  ``training=True`` is
        the reference's "noise" mode (:105-107): inputs + U(-1/2, 1/2) from a Philox stream keyed by ``seed`` (the reference
        is unseeded), likelihood evaluated at the noisy value."""
        x = runtime.unwrap(inputs)
        channels = x.shape[-1]
        c, slot = self._resolve(channels)
        xt = c.to_device(x, torch.float32)
        c.set_quantize_mode(bool(training), seed)
        try:
            x_hat, p, _, _ = c.factorized(slot, xt, self._likelihood_bound, want_p=True, want_bits=False)
        finally:
            c.set_quantize_mode(False)
        return runtime.DeviceResult(x_hat), runtime.DeviceResult(p)

    def estimate_bits(self, inputs) -> float:
        """sum(log2 p) * -1 over all elements (train_hyper.py:148-150 before the /num_points)."""
        x = runtime.unwrap(inputs)
        c, slot = self._resolve(x.shape[-1])
        _, _, bits, _ = c.factorized(slot, c.to_device(x, torch.float32), self._likelihood_bound, want_p=False, want_bits=True)
        return float(runtime.to_host(bits)[0])

    def _get_cdf(self, min_v, max_v):
        """int32 [1, C, N+1] like the reference (entropy_model.py:183-221)."""
        c, slot = self._resolve(self._codec.bn_channels[self._slot] if self._slot is not None else 8)
        return c.factorized_cdf(slot, int(min_v), int(max_v), self._likelihood_bound, self._range_coder_precision)[None]

    def compress_begin(self, inputs):
        """GPU half of ``compress``: quantise, global symbol range, per-channel CDF; returns what the host range coder
        needs so that the caller can run it on a worker thread."""
        x = runtime.unwrap(inputs)
        channels = x.shape[-1]
        c, slot = self._resolve(channels)
        xt = c.to_device(x, torch.float32)
        x_hat, _, _, mm = c.factorized(slot, xt, self._likelihood_bound, want_p=False, want_bits=False)
        mm_h = runtime.to_host(mm)
        min_v, max_v = int(mm_h[0]), int(mm_h[1])
        cdf = c.factorized_cdf(slot, min_v, max_v, self._likelihood_bound, self._range_coder_precision)
        sym = (runtime.to_host(x_hat).reshape(-1).astype(np.int32) - min_v).astype(np.int16)
        return sym, cdf, min_v, max_v

    def compress_quantized_host(self, x_hat: np.ndarray, chunk_minmax: np.ndarray):
        """Host-only tail of ``compress`` for latents that are already quantised and on the host (``x_hat`` float32 [..., C]) with
        the (min, max) of their parts (int32 [parts, 2]): global range (entropy_model.py:249-250), the per-channel CDF from the host
        twin of the CDF kernel (same det_math.h density and normaliser: bit-identical tables, tests/test_gpu_coder.py), ONE string.
        -> (bytes, min_v, max_v).  No GPU call: the pipeline runs it while the device is still busy."""
        channels = x_hat.shape[-1]
        c, slot = self._resolve(channels)
        mm = np.asarray(chunk_minmax).reshape(-1, 2)
        min_v, max_v = int(mm[:, 0].min()), int(mm[:, 1].max())
        cdf = c.factorized_cdf_host(slot, min_v, max_v, self._likelihood_bound, self._range_coder_precision)
        sym = (x_hat.reshape(-1).astype(np.int32) - min_v).astype(np.int16)
        return runtime.range_encode(sym, cdf, self._range_coder_precision), min_v, max_v

    def compress_finish(self, sym, cdf):
        return runtime.range_encode(sym, cdf, self._range_coder_precision)

    def compress(self, inputs):
        """-> (string, min_v, max_v): ONE string over the whole tensor with a global symbol range
        (entropy_model.py:223-261)."""
        sym, cdf, min_v, max_v = self.compress_begin(inputs)
        string = self.compress_finish(sym, cdf)
        return runtime.HostResult(string), runtime.HostResult(np.int32(min_v)), runtime.HostResult(np.int32(max_v))

    def decompress_progressive(self, strings, min_v, max_v, shape, channels=None):
        """As ``decompress`` but decoding on a worker thread: returns ``get(a, b)`` -> float32 device tensor of leading-axis
        slice [a, b) (blocks until those symbols are decoded).  The string is sequential, so the leading slices come first."""
        strings = runtime.unwrap(strings)
        if isinstance(strings, np.ndarray):
            strings = strings.item() if strings.ndim == 0 else strings.tobytes() if strings.dtype != object else strings[0]
        shape = [int(s) for s in np.asarray(runtime.unwrap(shape)).reshape(-1)]
        min_v, max_v = int(np.asarray(runtime.unwrap(min_v))), int(np.asarray(runtime.unwrap(max_v)))
        channels = int(np.asarray(runtime.unwrap(channels))) if channels is not None else shape[-1]
        c, slot = self._resolve(channels)
        cdf = c.factorized_cdf(slot, min_v, max_v, self._likelihood_bound, self._range_coder_precision)
        per = int(np.prod(shape[1:]))
        dec = runtime.ProgressiveDecode(strings, shape[0] * per, cdf, self._range_coder_precision)

        def get(a, b):
            sym = dec.wait(b * per)[a * per:b * per]
            vals = (sym.astype(np.int32) + min_v).astype(np.float32).reshape([b - a] + shape[1:])
            return c.to_device(vals)
        return get

    def decompress(self, strings, min_v, max_v, shape, channels=None):
        """-> float32 tensor of ``shape`` (entropy_model.py:263-306)."""
        strings = runtime.unwrap(strings)
        if isinstance(strings, np.ndarray):
            strings = strings.item() if strings.ndim == 0 else strings.tobytes() if strings.dtype != object else strings[0]
        shape = [int(s) for s in np.asarray(runtime.unwrap(shape)).reshape(-1)]
        min_v, max_v = int(np.asarray(runtime.unwrap(min_v))), int(np.asarray(runtime.unwrap(max_v)))
        channels = int(np.asarray(runtime.unwrap(channels))) if channels is not None else shape[-1]
        c, slot = self._resolve(channels)
        cdf = c.factorized_cdf(slot, min_v, max_v, self._likelihood_bound, self._range_coder_precision)
        n = int(np.prod(shape))
        sym = runtime.range_decode(strings, n, cdf, self._range_coder_precision)
        vals = (sym.astype(np.int32) + min_v).astype(np.float32).reshape(shape)
        return runtime.DeviceResult(c.to_device(vals))
