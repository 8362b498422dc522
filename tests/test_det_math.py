"""CPU-side checks of the pieces the GPU coder path shares with the host half of the library (no GPU needed):
the bit-reproducible likelihood math (csrc/det_math.h) through its host twin, and the Philox stream of the "noise" mode."""
import ctypes as C

import numpy as np

from oracle import coder as ocoder, entropy
from pcgcv1_b200 import _lib, runtime


def test_philox_known_answer_and_noise_hook():
    # Random123 known-answer vector: Philox4x32-10, counter 0, key 0
    r = entropy.philox4x32_10(np.array([0]), np.array([0]), 0, 0)
    assert [int(x[0]) for x in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = entropy.philox4x32_10(np.array([0xFFFFFFFF]), np.array([0xFFFFFFFF]), 0xFFFFFFFF, 0xFFFFFFFF)
    assert len(r) == 4
    L = _lib.lib()
    seed = 0x1234567890ABCDEF
    want = entropy.philox_uniform(seed, 64)
    out = np.zeros(4, np.float32)
    for v in range(16):
        assert L.pcgc_debug_noise(seed, v, out.ctypes.data) == 0
        assert np.array_equal(out, want[4 * v:4 * v + 4])
    assert want.min() >= -0.5 and want.max() < 0.5


def test_host_twin_rows_track_the_oracle():
    """pcgc_host_laplace_cdf (det_math.h exp, <= 2 ulp) against the oracle's rows (NumPy exp): valid tables, every entry within
    2 units, >= 99 % of the rows identical -- and the oracle decodes a stream coded with the twin's rows."""
    rng = np.random.default_rng(0)
    E = 8192
    loc = rng.normal(0, 3, E).astype(np.float32)
    scale = (np.abs(rng.normal(0, 1.5, E)) + 0.05).astype(np.float32)
    scale[:50] = 1e-9
    scale[50:100] = 100.0
    mm = np.array([[-12, 14]], np.int32)
    N = 27
    rows, off = runtime.host_laplace_cdf(loc[None], scale[None], mm, threads=2)
    assert off[-1] == E * N
    r = rows.reshape(E, N).astype(np.int64)
    assert (r[:, 0] == 0).all() and (np.diff(r, axis=1) >= 1).all() and (r[:, -1] < 65536).all()
    ref = entropy.SymmetricConditionalOracle().get_cdf(loc, scale, -12, 14)[:, :N]
    d = np.abs(r - ref)
    assert d.max() <= 2
    assert (d.max(axis=1) == 0).mean() >= 0.99
    # round trip: host encoder with the twin's rows -> oracle decoder with the same rows
    sym = rng.integers(0, N, E).astype(np.int16)
    table = np.concatenate([r, np.full((E, 1), 65536)], 1).astype(np.int32)
    lower = table[np.arange(E), sym].astype(np.uint32)
    width = (table[np.arange(E), sym + 1] - table[np.arange(E), sym] - 1).astype(np.uint32)
    iv = (lower | (width << 16)).astype(np.uint32)[None]
    s = runtime.range_encode_intervals_batch(iv, 1)[0]
    back = ocoder.range_decode(s, E, table, np.arange(E, dtype=np.int32), 16)
    assert np.array_equal(back, sym)
    host = runtime.range_decode_rows_batch([s], E, rows, off, mm, 1)
    assert np.array_equal(host[0], sym)
