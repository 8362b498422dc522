"""Oracle restatement of the reference's entropy models (test infrastructure only).

NumPy, computed in the dtype of the inputs (float32 = "O32", float64 = "O64").

* ``EntropyBottleneckOracle``  follows ``models/entropy_model.py``  (Balle-2018 factorized density
  with the reference's quirks: the tanh "factor" is applied on ALL four layers incl. the last,
  ``:86-96``; no median offsets, no tail mass; one GLOBAL min/max for the coder, ``:249-250``).
* ``SymmetricConditionalOracle`` follows ``models/conditional_entropy_model.py``  (the CDF is
  LAPLACE, ``:21-32``; the sign trick uses ``sign(upper+lower-loc)`` = ``sign(2x-loc)``, ``:47``).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import coder


def _softplus(x):
    # tf.nn.softplus: log(exp(x)+1), evaluated stably.
    return np.logaddexp(x, np.zeros_like(x))


def _sigmoid(x):
    # tf.math.sigmoid; evaluated stably for both signs.
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def philox4x32_10(c0, c1, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011, Random123) on uint32 arrays: counter (c0, c1, 0, 0), key (k0, k1) -> 4 uint32 arrays.
    The "noise" quantisation of the CUDA path draws from this generator (csrc/entropy.cu:noise4)."""
    c = [np.asarray(c0, np.uint64), np.asarray(c1, np.uint64), np.zeros_like(np.asarray(c0, np.uint64)), np.zeros_like(np.asarray(c0, np.uint64))]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return [x.astype(np.uint32) for x in c]


def philox_uniform(seed: int, n: int) -> np.ndarray:
    """n draws of U[-1/2, 1/2) as float32: element i uses word i % 4 of the block with counter i // 4, key = seed."""
    v = np.arange((n + 3) // 4, dtype=np.uint64)
    r = philox4x32_10(v & np.uint64(0xFFFFFFFF), v >> np.uint64(32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.stack(r, -1).reshape(-1)[:n]
    return ((u >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24) - np.float32(0.5)).astype(np.float32)


Y_NOISE_SEED_OFFSET = 0x9E3779B97F4A7C15      # the conditional model's stream (api.cu: pcgc_laplace_quantize_likelihood)


def tf_round(x):
    """tf.math.round = round half to even (np.rint)."""
    return np.rint(x)


class EntropyBottleneckOracle:
    """models/entropy_model.py:8-306."""

    def __init__(self, params: Dict[str, np.ndarray], likelihood_bound=1e-9, range_coder_precision=16):
        # params: matrix_i [C,f_{i+1},f_i], bais_i [C,f_{i+1},1], factor_i [C,f_{i+1},1], i=0..3
        self.n_layers = len([k for k in params if k.startswith("matrix_")])
        self.matrices = [np.asarray(params["matrix_%d" % i]) for i in range(self.n_layers)]
        self.biases = [np.asarray(params["bais_%d" % i]) for i in range(self.n_layers)]
        self.factors = [np.asarray(params["factor_%d" % i]) for i in range(self.n_layers)]
        self.channels = self.matrices[0].shape[0]
        self.likelihood_bound = likelihood_bound
        self.precision = range_coder_precision

    @staticmethod
    def default_params(channels: int, rng: np.random.Generator, init_scale=8.0, filters=(3, 3, 3)):
        """build(), entropy_model.py:25-70 (bias init U(-.5,.5); factor init zeros)."""
        f = (1,) + tuple(filters) + (1,)
        scale = init_scale ** (1.0 / (len(filters) + 1))
        p = {}
        for i in range(len(filters) + 1):
            init = np.log(np.expm1(1.0 / scale / f[i + 1]))
            p["matrix_%d" % i] = np.full((channels, f[i + 1], f[i]), init, np.float32)
            p["bais_%d" % i] = rng.uniform(-0.5, 0.5, (channels, f[i + 1], 1)).astype(np.float32)
            p["factor_%d" % i] = np.zeros((channels, f[i + 1], 1), np.float32)
        return p

    def logits_cumulative(self, inputs: np.ndarray) -> np.ndarray:
        """_logits_cumulative, entropy_model.py:72-98.  inputs: [C,1,M]."""
        dt = inputs.dtype
        logits = inputs
        for i in range(self.n_layers):
            m = _softplus(self.matrices[i].astype(dt))
            logits = np.matmul(m, logits)
            logits = logits + self.biases[i].astype(dt)
            f = np.tanh(self.factors[i].astype(dt))
            logits = logits + f * np.tanh(logits)
        return logits

    def _likelihood_cm(self, x_cm: np.ndarray) -> np.ndarray:
        dt = x_cm.dtype
        half = dt.type(0.5)
        lower = self.logits_cumulative(x_cm - half)
        upper = self.logits_cumulative(x_cm + half)
        sign = -np.sign(lower + upper)
        return np.abs(_sigmoid(sign * upper) - _sigmoid(sign * lower))

    def likelihood(self, x: np.ndarray) -> np.ndarray:
        """_likelihood, entropy_model.py:114-151 (channels-last in, channels-last out)."""
        c = x.shape[-1]
        x_cm = np.moveaxis(x, -1, 0).reshape(c, 1, -1)
        p = self._likelihood_cm(x_cm)
        return np.moveaxis(p.reshape((c,) + x.shape[:-1]), 0, -1)

    def __call__(self, x: np.ndarray, training: bool = False, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """call(inputs, training), entropy_model.py:153-181; training=True is the "noise" mode (:105-107)."""
        x_hat = (x + philox_uniform(seed, x.size).reshape(x.shape).astype(x.dtype)) if training else tf_round(x)
        p = np.maximum(self.likelihood(x_hat), x.dtype.type(self.likelihood_bound))
        return x_hat, p

    def pmf(self, min_v: int, max_v: int, dtype=np.float32) -> np.ndarray:
        """_get_cdf up to the pmf, entropy_model.py:183-215.  Returns [C, N]."""
        a = np.arange(min_v, max_v + 1, dtype=dtype).reshape(1, 1, -1)
        a = np.tile(a, (self.channels, 1, 1))
        p = np.maximum(self._likelihood_cm(a), dtype(self.likelihood_bound))
        return p[:, 0, :]

    def get_cdf(self, min_v: int, max_v: int) -> np.ndarray:
        """_get_cdf, entropy_model.py:183-221: int32 [C, N+1]."""
        return coder.pmf_to_quantized_cdf(self.pmf(min_v, max_v), self.precision)

    def compress(self, x: np.ndarray):
        """compress, entropy_model.py:223-261: ONE string over all elements, global min/max."""
        values = tf_round(x.astype(np.float32))
        min_v = int(np.floor(values.min()))
        max_v = int(np.ceil(values.max()))
        cdf = self.get_cdf(min_v, max_v)                           # [C, N+1]
        sym = (values.reshape(-1, self.channels).astype(np.int32) - min_v).astype(np.int16)
        idx = np.tile(np.arange(self.channels, dtype=np.int32), sym.shape[0])
        return coder.range_encode(sym.reshape(-1), cdf, idx, self.precision), min_v, max_v

    def decompress(self, string: bytes, min_v: int, max_v: int, shape):
        """decompress, entropy_model.py:263-306."""
        cdf = self.get_cdf(int(min_v), int(max_v))
        n = int(np.prod(shape))
        idx = np.tile(np.arange(self.channels, dtype=np.int32), n // self.channels)
        sym = coder.range_decode(string, n, cdf, idx, self.precision)
        return (sym.astype(np.int32) + int(min_v)).reshape(shape).astype(np.float32)


class SymmetricConditionalOracle:
    """models/conditional_entropy_model.py:8-201."""

    def __init__(self, likelihood_bound=1e-9, range_coder_precision=16):
        self.likelihood_bound = likelihood_bound
        self.precision = range_coder_precision

    @staticmethod
    def standardized_cumulative(t, loc, scale):
        """_standardized_cumulative (Laplace), conditional_entropy_model.py:21-32."""
        dt = np.result_type(t, loc, scale)
        half = dt.type(0.5)
        one = dt.type(1.0)
        e = np.exp(-np.abs(t - loc) / scale)
        c_l = half * e
        c_r = one - half * e
        return c_l * (t <= loc).astype(dt) + c_r * (t > loc).astype(dt)

    def likelihood(self, x, loc, scale):
        """_likelihood, conditional_entropy_model.py:34-56."""
        dt = np.result_type(x, loc, scale)
        half = dt.type(0.5)
        upper = x + half
        lower = x - half
        sign = np.sign(upper + lower - loc)
        upper = -sign * (upper - loc) + loc
        lower = -sign * (lower - loc) + loc
        return np.abs(self.standardized_cumulative(upper, loc, scale)
                      - self.standardized_cumulative(lower, loc, scale))

    def __call__(self, y, loc, scale, training: bool = False, seed: int = 0):
        """call(inputs, loc, scale, training), conditional_entropy_model.py:71-93; training=True = "noise" (:62-64)."""
        y_hat = (y + philox_uniform((seed + Y_NOISE_SEED_OFFSET) & 0xFFFFFFFFFFFFFFFF, y.size).reshape(y.shape).astype(y.dtype)) if training else tf_round(y)
        p = np.maximum(self.likelihood(y_hat, loc, scale), y.dtype.type(self.likelihood_bound))
        return y_hat, p

    def pmf(self, loc, scale, min_v: int, max_v: int):
        """_get_cdf up to the pmf, conditional_entropy_model.py:95-120.  loc/scale [R] -> [R,N]."""
        dt = loc.dtype
        a = np.arange(min_v, max_v + 1, dtype=dt)[None, :]
        p = self.likelihood(a, loc.reshape(-1, 1), scale.reshape(-1, 1))
        return np.maximum(p, dt.type(self.likelihood_bound))

    def get_cdf(self, loc, scale, min_v: int, max_v: int) -> np.ndarray:
        """_get_cdf, conditional_entropy_model.py:95-124: int32 [R, N+1], one row PER ELEMENT."""
        return coder.pmf_to_quantized_cdf(self.pmf(loc, scale, min_v, max_v), self.precision)

    def compress(self, y, loc, scale):
        """compress, conditional_entropy_model.py:126-163 (caller passes ONE cube)."""
        values = tf_round(y.astype(np.float32)).reshape(-1)
        min_v = int(np.floor(values.min()))
        max_v = int(np.ceil(values.max()))
        cdf = self.get_cdf(loc.reshape(-1).astype(np.float32), scale.reshape(-1).astype(np.float32), min_v, max_v)
        sym = (values.astype(np.int32) - min_v).astype(np.int16)
        idx = np.arange(sym.size, dtype=np.int32)
        return coder.range_encode(sym, cdf, idx, self.precision), min_v, max_v

    def decompress(self, string: bytes, loc, scale, min_v: int, max_v: int, shape):
        """decompress, conditional_entropy_model.py:165-201."""
        cdf = self.get_cdf(loc.reshape(-1).astype(np.float32), scale.reshape(-1).astype(np.float32), int(min_v), int(max_v))
        n = int(np.prod(shape))
        idx = np.arange(n, dtype=np.int32)
        sym = coder.range_decode(string, n, cdf, idx, self.precision)
        return (sym.astype(np.int32) + int(min_v)).reshape(shape).astype(np.float32)


def estimated_bits(p: np.ndarray) -> float:
    """sum(log p) / -ln 2, train_hyper.py:148-150 (before the division by num_points)."""
    return float(np.sum(np.log(p.astype(np.float64))) / -np.log(2.0))
