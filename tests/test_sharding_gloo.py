"""Host logic of the multi-GPU path on CPU: world_size-2 gloo, per-rank work done by a stub
codec built from the ORACLE (tests may use it), compared with the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcgcv1_b200 import sharding


def test_shard_slices():
    assert sharding.shard_slices(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.shard_slices(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharding.shard_slices(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 191, 7769):
        for w in (1, 2, 4, 8):
            s = sharding.shard_slices(n, w)
            assert s[0][0] == 0 and s[-1][1] == n and all(a[1] == b[0] for a, b in zip(s, s[1:]))
            sizes = [b - a for a, b in s]
            assert max(sizes) - min(sizes) <= 1


def test_decode_slices_cover_in_order_and_equalise_the_finish_times():
    for n in (0, 1, 5, 7, 191, 7769):
        for w in (1, 2, 4, 8):
            for ratio in (0.5, 2.2, 3.0, 10.0):
                s = sharding.decode_slices(n, w, ratio)
                assert len(s) == w and s[0][0] == 0 and s[-1][1] == n and all(a[1] == b[0] for a, b in zip(s, s[1:]))
                assert all(b >= a for a, b in s)
                sizes = [b - a for a, b in s]
                assert sizes == sorted(sizes, reverse=True) or n < 4 * w        # earlier ranks wait less for the string: more cubes
    assert sharding.decode_slices(100, 1, 3.0) == [(0, 100)]
    assert sharding.decode_slices(100, 4, 0.0) == sharding.shard_slices(100, 4)          # <= 0: the balanced slices
    assert sharding.decode_slices(100, 4, -1.0) == sharding.shard_slices(100, 4)
    # the model behind it: rank 0 works through its slice at the GPU rate, rank r > 0 starts when the string decoder (ratio x the GPU
    # rate) has passed the end of its slice -> every rank finishes at the same time (to within the rounding to whole cubes)
    n, w, ratio = 7769, 8, 3.0
    s = sharding.decode_slices(n, w, ratio)
    finish = [(b - a) if r == 0 else b / ratio + (b - a) for r, (a, b) in enumerate(s)]
    assert max(finish) - min(finish) <= 2.0
    balanced = sharding.shard_slices(n, w)
    assert max(finish) < 0.85 * max((b - a) if r == 0 else b / ratio + (b - a) for r, (a, b) in enumerate(balanced))
    big = sharding.decode_slices(n, w, 1e9)                                       # an instant string decoder: balanced to within a cube
    assert max(abs((b - a) - n / w) for a, b in big) <= 1.0


class OracleLocalCodec(sharding.LocalCodec):
    """CPU stand-in: tiny 'latents' derived from the cubes, real oracle entropy coding."""

    def __init__(self):
        from oracle import entropy
        from pcgcv1_b200 import weights as W
        self.eb = entropy.EntropyBottleneckOracle(W.entropy_bottleneck_params(8, np.random.default_rng(5)))
        self.sc = entropy.SymmetricConditionalOracle()

    @staticmethod
    def _latents(cube):
        rng = np.random.default_rng(int(cube.sum()) + 1)
        y = rng.normal(0, 2, (1, 4, 4, 4, 16)).astype(np.float32)
        z = rng.normal(0, 2, (1, 8, 8, 8, 8)).astype(np.float32)
        return y, z

    @staticmethod
    def _params(z_hat):
        loc = np.resize(z_hat.astype(np.float32), (1, 4, 4, 4, 16)) * 0.5
        scale = np.abs(loc) * 0.3 + 0.5
        return loc, scale

    def encode_local(self, cubes):
        out = {"y_strings": [], "y_min": [], "y_max": [], "z_hat": []}
        for c in cubes:
            y, z = self._latents(c)
            z_hat = np.rint(z)
            loc, scale = self._params(z_hat)
            s, mn, mx = self.sc.compress(y, loc, scale)
            out["y_strings"].append(s); out["y_min"].append(mn); out["y_max"].append(mx); out["z_hat"].append(z_hat[0].astype(np.int16))
        out["y_min"] = np.array(out["y_min"], np.int32); out["y_max"] = np.array(out["y_max"], np.int32)
        out["z_hat"] = np.array(out["z_hat"], np.int16).reshape(-1, 8, 8, 8, 8)
        return out

    def encode_z(self, z_hat_all):
        return self.eb.compress(z_hat_all.astype(np.float32))

    def decode_z(self, z_string, z_min, z_max, z_shape):
        return self.eb.decompress(z_string, z_min, z_max, tuple(int(v) for v in z_shape)).astype(np.int16)

    def decode_local(self, y_strings, y_min, y_max, z_hat, nums, rho):
        masks = np.zeros((len(y_strings), 4, 4, 4, 16), np.uint8)
        for i, s in enumerate(y_strings):
            loc, scale = self._params(z_hat[i:i + 1].astype(np.float32))
            y_hat = self.sc.decompress(s, loc, scale, int(y_min[i]), int(y_max[i]), (1, 4, 4, 4, 16))
            masks[i] = (y_hat[0] > 0).astype(np.uint8)
        return masks


def _cubes(n):
    rng = np.random.default_rng(0)
    return (rng.random((n, 8, 8, 8, 1)) < 0.1).astype(np.uint8)


def _worker(rank, world, port, n_cubes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cubes = _cubes(n_cubes)
        a, b = sharding.shard_slices(n_cubes, world)[rank]
        local = OracleLocalCodec()
        stream = sharding.compress_sharded(cubes[a:b], local)
        masks = sharding.decompress_sharded(stream, np.arange(n_cubes) if rank == 0 else None, 1.0, local)
        points = sharding.decompress_sharded(stream, np.arange(n_cubes) if rank == 0 else None, 1.0, local, output="points")
        if rank == 0:
            q.put((stream, masks, points))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cubes", [5, 1])
def test_world2_matches_single_process(n_cubes):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cubes, q)) for r in range(2)]
    for p in procs:
        p.start()
    stream, masks, points = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    local = OracleLocalCodec()
    part = local.encode_local(_cubes(n_cubes))
    z_string, z_min, z_max = local.encode_z(part["z_hat"])
    assert stream["y_strings"] == part["y_strings"]
    assert np.array_equal(stream["y_min"], part["y_min"]) and np.array_equal(stream["y_max"], part["y_max"])
    assert stream["z_string"] == z_string and (stream["z_min"], stream["z_max"]) == (z_min, z_max)
    ref = local.decode_local(part["y_strings"], part["y_min"], part["y_max"], part["z_hat"], np.arange(n_cubes), 1.0)
    assert np.array_equal(masks, ref)
    # output="points": the same voxels as coordinates, cube by cube in order
    pts, counts = points
    assert np.array_equal(counts, ref.reshape(n_cubes, -1).sum(1))
    assert np.array_equal(pts, np.concatenate([np.argwhere(m > 0) for m in ref]).astype(np.int16))
