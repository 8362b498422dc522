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
  renormalisation when ``size-1 < 2^16``, big-endian 16-bit words.  ``UpstreamRangeEncoder`` restates the
  upstream ``RangeEncoder::Encode`` / ``Finalize`` state machine LITERALLY (32-bit base that may wrap, the
  ``delay_`` word + byte counter for an interval that straddles 2^32, Finalize = "2^32" in the delayed
  state, else base rounded up to a multiple of 2^16 with a zero low byte left out and nothing for base 0).
  ``_Encoder`` is the carry-propagating form the product's C++ / CUDA coders use (33-bit base, delayed
  word + counter of pending 0xFFFF words) with the same Finalize rule; the two agree byte for byte on
  every input (tests/test_host_coder.py), so the product is pinned on the upstream ALGORITHM -- a TF
  binary to pin the restatement itself is what is missing.

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
        """RangeEncoder::Finalize of the upstream coder, in this state machine's terms.  The interval is [base, base + size_minus1]
        (base exact, up to 33 bits).  If it still holds a multiple of 2^32 above base (upstream: ``delay_ != 0``), that multiple is
        the value written: the delayed word + 1, everything after it zero and left out.  Otherwise base is rounded up to a multiple
        of 2^16 (upstream ``mid``): the words before the last one go out in full, of the last one the low byte only if it is not
        zero, and nothing at all if the low 32 bits of base are zero."""
        carry, low32 = self.base >> 32, self.base & 0xFFFFFFFF
        if carry == 0 and low32 + self.size_minus1 > 0xFFFFFFFF:
            w = (self.cache + 1) & 0xFFFF           # a straddling interval only exists after a renormalisation: cache is set
            self.out.append(w >> 8)
            if w & 0xFF:
                self.out.append(w & 0xFF)
            return bytes(self.out)
        if self.cache is not None:
            self._emit16((self.cache + carry) & 0xFFFF)
        for _ in range(self.pending):
            self._emit16((0xFFFF + carry) & 0xFFFF)
        if low32 != 0:
            mid = ((low32 - 1) >> 16) + 1
            self.out.append(mid >> 8)
            if mid & 0xFF:
                self.out.append(mid & 0xFF)
        return bytes(self.out)


class UpstreamRangeEncoder:
    """``RangeEncoder`` of tensorflow/contrib/coder/kernels/range_coder.cc (tensorflow-gpu==1.13.1), restated statement by
    statement: ``Encode(lower, upper)`` and ``Finalize()``.  uint32 ``base_`` / ``size_minus1_`` wrap like the C++ types; ``delay_``
    holds, in its low 16 bits, the delayed word + 1 and, above them, the number of delayed BYTES after it."""

    M = 0xFFFFFFFF

    def __init__(self, precision: int):
        self.precision = precision
        self.base = 0
        self.size_minus1 = self.M
        self.delay = 0
        self.out = bytearray()

    def encode(self, lower: int, upper: int):
        M = self.M
        size = self.size_minus1 + 1
        a = ((size * lower) >> self.precision) & M
        b = (((size * upper) >> self.precision) - 1) & M
        self.base = (self.base + a) & M
        self.size_minus1 = (b - a) & M
        base_overflow = self.base < a
        if ((self.base + self.size_minus1) & M) < self.base:
            # the interval [base, base + size) wraps around 2^32: the carry is not known yet
            assert self.delay & 0xFFFF
            if (self.size_minus1 >> 16) == 0:
                assert (self.base >> 16) == 0xFFFF
                self.base = (self.base << 16) & M
                self.size_minus1 = ((self.size_minus1 << 16) | 0xFFFF) & M
                self.delay += 0x20000                        # two more delayed bytes
            return
        if self.delay != 0:
            if base_overflow:                                # carry: delayed word + 1, then zero bytes
                self.out.append((self.delay >> 8) & 0xFF)
                self.out.append(self.delay & 0xFF)
                self.out.extend(b"\x00" * (self.delay >> 16))
            else:                                            # no carry: the delayed word, then 0xFF bytes
                self.delay -= 1
                self.out.append((self.delay >> 8) & 0xFF)
                self.out.append(self.delay & 0xFF)
                self.out.extend(b"\xff" * (self.delay >> 16))
            self.delay = 0
        if (self.size_minus1 >> 16) == 0:
            top = self.base >> 16
            self.base = (self.base << 16) & M
            self.size_minus1 = ((self.size_minus1 << 16) | 0xFFFF) & M
            if self.base <= ((self.base + self.size_minus1) & M):
                self.out.append((top >> 8) & 0xFF)
                self.out.append(top & 0xFF)
            else:
                assert top < 0xFFFF
                self.delay = top + 1

    def finish(self) -> bytes:
        if self.delay != 0:                                  # the last state was the wrapped one: take the value 2^32
            self.out.append((self.delay >> 8) & 0xFF)
            if self.delay & 0xFF:
                self.out.append(self.delay & 0xFF)
        elif self.base != 0:                                 # base rounded up to the next multiple of 2^16 (base 0: nothing)
            mid = ((self.base - 1) >> 16) + 1
            assert mid & 0xFFFF == mid
            self.out.append((mid >> 8) & 0xFF)
            if mid & 0xFF:
                self.out.append(mid & 0xFF)
        return bytes(self.out)


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
