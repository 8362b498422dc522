"""Host-side plumbing of the coder pipeline that needs no GPU: chunk schedule, the progress-publishing hyper-string decoder
and the multithreaded result copy."""
import numpy as np

from pcgcv1_b200 import runtime, transform


def test_chunk_schedule_covers_every_cube_once():
    for B in (0, 1, 15, 16, 31, 32, 33, 64, 65, 191, 1000, 7769):
        for kw in ({}, {"small_first": True}, {"small_last": True}, {"small_first": True, "small_last": True}):
            ch = transform._chunks(B, **kw)
            assert sum(b - a for a, b in ch) == B
            assert all(a < b for a, b in ch)
            assert all(ch[i][1] == ch[i + 1][0] for i in range(len(ch) - 1))
            if ch:
                assert ch[0][0] == 0 and ch[-1][1] == B and max(b - a for a, b in ch) <= transform._CHUNK
    assert transform._chunks(191, small_first=True)[0] == (0, transform._CHUNK_EDGE)
    assert transform._chunks(191, small_last=True)[-1] == (191 - transform._CHUNK_EDGE, 191)


def _cdf(C, N, rng):
    cdf = np.zeros((C, N + 1), np.int32)
    for c in range(C):
        q = rng.integers(1, 2000, size=N).astype(np.int64)
        q = np.maximum(1, q * 65536 // q.sum())
        q[int(np.argmax(q))] += 65536 - q.sum()
        cdf[c, 1:] = np.cumsum(q)
    return cdf


def test_progressive_decode_equals_one_shot_decode():
    rng = np.random.default_rng(5)
    C, N, n = 8, 19, 8 * 30000
    cdf = _cdf(C, N, rng)
    sym = rng.integers(0, N, size=n).astype(np.int16)
    data = runtime.range_encode(sym, cdf)
    want = runtime.range_decode(data, n, cdf)
    assert np.array_equal(want, sym)
    dec = runtime.ProgressiveDecode(data, n, cdf, step=4096)
    head = dec.wait(10000).copy()
    assert np.array_equal(head, sym[:10000])
    assert np.array_equal(dec.wait(n + 5), sym)                 # clamped to n
    assert np.array_equal(dec.finish(), sym)
    empty = runtime.ProgressiveDecode(b"", 0, cdf)
    assert empty.wait(0).shape == (0,) and empty.finish().shape == (0,)


def test_host_copy_small_and_large():
    rng = np.random.default_rng(6)
    for n in (0, 1, 1000, (8 << 20) + 12345, 40_000_001):
        a = rng.integers(0, 255, size=n, dtype=np.uint8)
        b = runtime.host_copy(a)
        assert b is not a and b.dtype == a.dtype and np.array_equal(a, b)
    m = rng.integers(0, 2, size=(3, 64, 64, 64, 1), dtype=np.uint8)
    assert np.array_equal(runtime.host_copy(m[:, ::2]), m[:, ::2])          # non-contiguous input
