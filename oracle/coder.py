"""Oracle restatement of ``tensorflow.contrib.coder`` ops (test infrastructure only).

**parity unpinned**: the reference calls ``coder_ops.pmf_to_quantized_cdf / range_encode /
range_decode`` (``models/entropy_model.py:218,258,298``; ``models/conditional_entropy_model.py:
122,161,195``) from the un-vendored wheel ``tensorflow-gpu==1.13.1``
(``tensorflow/contrib/coder/kernels/{pmf_to_cdf_op,range_coder,range_coder_ops}.cc``).  Neither
the wheel nor any golden bitstream exists offline, so this file restates the PUBLISHED
algorithm and the tests anchor on round trips, CDF validity and coded size:

* ``pmf_to_quantized_cdf(pmf, precision)``: per row ``v_i = max(rint(pmf_i * 2^precision), 1)``
  (no renormalisation of the pmf); while ``sum(v) > 2^precision`` decrement the entry with the
  smallest penalty ``pmf_i * (log2 v_i - log2(v_i - 1))`` (entries at 1 are never decremented);
  while ``sum(v) < 2^precision`` increment the entry with the largest gain
  ``pmf_i * (log2(v_i + 1) - log2 v_i)``; ``cdf = [0, cumsum(v)]`` (int32).  Ties: lowest index
  (the upstream tie order is an artefact of ``std::sort`` and is not specified).
* range coder: 32-bit ``base`` / ``size-1`` state, interval update
  ``a = (size*lower) >> precision``, ``b = ((size*upper) >> precision) - 1``, 16-bit
  renormalisation when ``size-1 < 2^16``, carries propagated through a delayed word + a
  counter of pending 0xFFFF words, big-endian 16-bit words, finalisation picks the multiple
  of 2^16 inside the interval and drops trailing zero bytes (the decoder pads zeros).

The pure-Python versions below are the definition; ``oracle/c/oracle_coder.c`` is the same
algorithm in C (built by ``oracle/build.py``) and is cross-checked against them in the tests.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB: Optional[ctypes.CDLL] = None


def _clib() -> Optional[ctypes.CDLL]:
    """Load the C restatement if it has been built (oracle/build.py)."""
    global _CLIB
    if _CLIB is None:
        path = os.path.join(_HERE, "liboracle_coder.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.orc_pmf_to_quantized_cdf.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            lib.orc_pmf_to_quantized_cdf.restype = ctypes.c_int
            lib.orc_range_encode.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
            lib.orc_range_encode.restype = ctypes.c_int64
            lib.orc_range_decode.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
            lib.orc_range_decode.restype = ctypes.c_int
            _CLIB = lib
    return _CLIB


# ----------------------------------------------------------------------------- pmf -> cdf
def _pmf_row_to_cdf_py(pmf_row: np.ndarray, precision: int) -> np.ndarray:
    target = 1 << precision
    mass = pmf_row.astype(np.float64)
    v = np.maximum(np.rint(pmf_row.astype(np.float32) * np.float32(target)).astype(np.int64), 1)
    s = int(v.sum())
    while s > target:
        pen = np.where(v > 1, mass * (np.log2(v) - np.log2(np.maximum(v - 1, 1))), np.inf)
        i = int(np.argmin(pen))          # first (lowest-index) minimum
        if not np.isfinite(pen[i]):
            raise ValueError("pmf_to_quantized_cdf: cannot shrink a row of ones")
        v[i] -= 1
        s -= 1
    while s < target:
        gain = mass * (np.log2(v + 1) - np.log2(v))
        i = int(np.argmax(gain))         # first (lowest-index) maximum
        v[i] += 1
        s += 1
    out = np.zeros(v.size + 1, np.int32)
    out[1:] = np.cumsum(v)
    return out


def pmf_to_quantized_cdf(pmf: np.ndarray, precision: int = 16, force_python: bool = False) -> np.ndarray:
    """float32 pmf [..., N] -> int32 cdf [..., N+1] (N must be > 1, as upstream requires)."""
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    n = pmf.shape[-1]
    if n < 2:
        raise ValueError("pmf_to_quantized_cdf: `pmf` size should be at least 2 in the last axis")
    rows = pmf.reshape(-1, n)
    out = np.empty((rows.shape[0], n + 1), np.int32)
    lib = None if force_python else _clib()
    if lib is not None:
        rc = lib.orc_pmf_to_quantized_cdf(rows.ctypes.data, rows.shape[0], n, precision, out.ctypes.data)
        if rc != 0:
            raise ValueError("pmf_to_quantized_cdf failed (rc=%d)" % rc)
    else:
        for r in range(rows.shape[0]):
            out[r] = _pmf_row_to_cdf_py(rows[r], precision)
    return out.reshape(pmf.shape[:-1] + (n + 1,))


# ----------------------------------------------------------------------------- range coder
class _Encoder:
    def __init__(self, precision: int):
        self.precision = precision
        self.base = 0                 # up to 33 bits: bit 32 is a not-yet-propagated carry
        self.size_minus1 = 0xFFFFFFFF
        self.cache = None             # delayed 16-bit word (None until the first shift)
        self.pending = 0              # number of delayed 0xFFFF words after `cache`
        self.out = bytearray()

    def _emit16(self, w: int):
        self.out.append((w >> 8) & 0xFF)
        self.out.append(w & 0xFF)

    def _shift(self):
        carry = self.base >> 32
        low32 = self.base & 0xFFFFFFFF
        if low32 < 0xFFFF0000 or carry:
            if self.cache is not None:
                self._emit16((self.cache + carry) & 0xFFFF)
            for _ in range(self.pending):
                self._emit16((0xFFFF + carry) & 0xFFFF)
            self.pending = 0
            self.cache = (low32 >> 16) & 0xFFFF
        else:
            self.pending += 1
        self.base = (low32 & 0xFFFF) << 16

    def encode(self, lower: int, upper: int):
        assert 0 <= lower < upper <= (1 << self.precision)
        size = self.size_minus1 + 1
        a = (size * lower) >> self.precision
        b = ((size * upper) >> self.precision) - 1
        self.base += a
        self.size_minus1 = b - a
        if (self.size_minus1 >> 16) == 0:
            self._shift()
            self.size_minus1 = ((self.size_minus1 << 16) | 0xFFFF) & 0xFFFFFFFF

    def finish(self) -> bytes:
        v = (self.base + 0xFFFF) >> 16          # round base up to a multiple of 2^16 (17-bit)
        carry, word = v >> 16, v & 0xFFFF
        if self.cache is not None:
            self._emit16((self.cache + carry) & 0xFFFF)
        for _ in range(self.pending):
            self._emit16((0xFFFF + carry) & 0xFFFF)
        self._emit16(word)
        out = self.out
        n = len(out)
        while n > 0 and out[n - 1] == 0:
            n -= 1
        return bytes(out[:n])


class _Decoder:
    def __init__(self, data: bytes, precision: int):
        self.data = data
        self.pos = 0
        self.precision = precision
        self.base = 0
        self.size_minus1 = 0xFFFFFFFF
        self.value = (self._read16() << 16) | self._read16()

    def _read16(self) -> int:
        v = 0
        for _ in range(2):
            v <<= 8
            if self.pos < len(self.data):
                v |= self.data[self.pos]
                self.pos += 1
        return v

    def decode(self, cdf_row) -> int:
        size = self.size_minus1 + 1
        offset = ((((self.value - self.base) & 0xFFFFFFFF) + 1) << self.precision) - 1
        lo, hi = 1, len(cdf_row) - 1            # first i in [1, N] with size*cdf[i] > offset
        while lo < hi:
            mid = (lo + hi) // 2
            if size * int(cdf_row[mid]) > offset:
                hi = mid
            else:
                lo = mid + 1
        s = lo - 1
        a = (size * int(cdf_row[s])) >> self.precision
        b = ((size * int(cdf_row[s + 1])) >> self.precision) - 1
        self.base = (self.base + a) & 0xFFFFFFFF
        self.size_minus1 = b - a
        if (self.size_minus1 >> 16) == 0:
            self.base = (self.base << 16) & 0xFFFFFFFF
            self.size_minus1 = ((self.size_minus1 << 16) | 0xFFFF) & 0xFFFFFFFF
            self.value = ((self.value << 16) & 0xFFFFFFFF) | self._read16()
        return s


def range_encode(symbols: np.ndarray, cdf: np.ndarray, cdf_index: np.ndarray, precision: int = 16,
                 force_python: bool = False) -> bytes:
    """Encode ``symbols[i]`` (int16, row-major order) with CDF row ``cdf[cdf_index[i]]``.

    ``cdf`` is int32 ``[rows, N+1]``; the reference's broadcasting of the leading CDF
    dimensions against the data (range_coder_ops.cc) is expressed by ``cdf_index``."""
    symbols = np.ascontiguousarray(symbols, dtype=np.int16).reshape(-1)
    cdf = np.ascontiguousarray(cdf, dtype=np.int32)
    cdf_index = np.ascontiguousarray(cdf_index, dtype=np.int32).reshape(-1)
    lib = None if force_python else _clib()
    if lib is not None:
        cap = 4 * symbols.size + 16
        buf = np.empty(cap, np.uint8)
        n = lib.orc_range_encode(symbols.ctypes.data, symbols.size, cdf.ctypes.data, cdf.shape[1],
                                 cdf_index.ctypes.data, precision, buf.ctypes.data, cap)
        if n < 0:
            raise ValueError("range_encode failed (rc=%d)" % n)
        return buf[:n].tobytes()
    enc = _Encoder(precision)
    nsym = cdf.shape[1] - 1
    for s, r in zip(symbols.tolist(), cdf_index.tolist()):
        if not 0 <= s < nsym:
            raise ValueError("symbol out of range")
        enc.encode(int(cdf[r, s]), int(cdf[r, s + 1]))
    return enc.finish()


def range_decode(data: bytes, count: int, cdf: np.ndarray, cdf_index: np.ndarray, precision: int = 16,
                 force_python: bool = False) -> np.ndarray:
    cdf = np.ascontiguousarray(cdf, dtype=np.int32)
    cdf_index = np.ascontiguousarray(cdf_index, dtype=np.int32).reshape(-1)
    lib = None if force_python else _clib()
    if lib is not None:
        out = np.empty(count, np.int16)
        src = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
        rc = lib.orc_range_decode(src.ctypes.data, len(data), count, cdf.ctypes.data, cdf.shape[1],
                                  cdf_index.ctypes.data, precision, out.ctypes.data)
        if rc != 0:
            raise ValueError("range_decode failed (rc=%d)" % rc)
        return out
    dec = _Decoder(bytes(data), precision)
    out = np.empty(count, np.int16)
    idx = cdf_index.tolist()
    for i in range(count):
        out[i] = dec.decode(cdf[idx[i]])
    return out
