"""Multi-GPU form of the codec: independent cubes are partitioned across ranks, no data-path collective.

SURVEY.md section 8(e): every cube is analysed, hyper-coded, entropy-modelled, synthesised and classified
independently (transform.py:42-48,116-168,238-256), so rank r owns the contiguous slice
``[r*B/W, (r+1)*B/W)`` of the reference's cube order and weights are replicated.  The only exchange is a
host-side ordered gather of per-cube results plus ONE global step: the hyper-latents z of all cubes form
a single string with a single (min_v, max_v) (models/entropy_model.py:249-259), so rank 0 range-codes z
after gathering the quantised z (4 KiB per cube).

``torch.distributed`` (NCCL group on GPUs, gloo in the CPU tests) is used only for ``gather_object`` /
``scatter_object_list`` of small host objects.  The per-rank work is behind the ``LocalCodec`` protocol so
the host logic is testable without a GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_slices(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, count-balanced slices; the first ``n % world`` ranks get one extra cube."""
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


class LocalCodec:
    """What a rank must provide (see ``GpuLocalCodec``)."""

    def encode_local(self, cubes: np.ndarray) -> dict:
        """-> {'y_strings': [bytes], 'y_min': int32[b], 'y_max': int32[b], 'z_hat': int array [b,8,8,8,8]}"""
        raise NotImplementedError

    def encode_z(self, z_hat_all: np.ndarray) -> Tuple[bytes, int, int]:
        raise NotImplementedError

    def decode_z(self, z_string: bytes, z_min: int, z_max: int, z_shape) -> np.ndarray:
        raise NotImplementedError

    def decode_local(self, y_strings: Sequence[bytes], y_min, y_max, z_hat: np.ndarray, nums: np.ndarray, rho: float) -> np.ndarray:
        """-> uint8 occupancy masks [b,64,64,64,1] of the top int(rho*nums) voxels"""
        raise NotImplementedError


class GpuLocalCodec(LocalCodec):
    """The CUDA implementation: one ``runtime.Codec`` on this rank's GPU."""

    def __init__(self, model="voxception", ckpt_dir="", device: Optional[int] = None):
        from . import runtime
        from .models.conditional_entropy_model import SymmetricConditional
        from .models.entropy_model import EntropyBottleneck
        self.runtime = runtime
        self.codec = runtime.get_codec(model, ckpt_dir, device)
        self.eb = EntropyBottleneck().bind(self.codec, self.codec.bottleneck_slot(8))
        self.sc = SymmetricConditional().bind(self.codec)

    def encode_local(self, cubes):
        import torch
        c = self.codec
        if len(cubes) == 0:
            return {"y_strings": [], "y_min": np.zeros(0, np.int32), "y_max": np.zeros(0, np.int32), "z_hat": np.zeros((0, 8, 8, 8, 8), np.int16)}
        ys = c.analysis(c.to_device(cubes))
        z_hat, _, _, _ = c.factorized(self.eb._slot, c.hyper_encode(ys), want_p=False, want_bits=False)
        locs, scales = c.hyper_decode(z_hat, 1e-9)
        strings, mn, mx = self.sc.compress_cubes(ys, locs, scales)
        return {"y_strings": strings, "y_min": mn.astype(np.int32), "y_max": mx.astype(np.int32),
                "z_hat": self.runtime.to_host(z_hat).astype(np.int16)}

    def encode_z(self, z_hat_all):
        s, mn, mx = self.eb.compress(z_hat_all.astype(np.float32))
        return s.numpy(), int(mn), int(mx)

    def decode_z(self, z_string, z_min, z_max, z_shape):
        return self.eb.decompress(z_string, z_min, z_max, np.asarray(z_shape), z_shape[-1]).numpy().astype(np.int16)

    def decode_local(self, y_strings, y_min, y_max, z_hat, nums, rho):
        import torch
        c = self.codec
        if len(y_strings) == 0:
            return np.zeros((0, 64, 64, 64, 1), np.uint8)
        locs, scales = c.hyper_decode(c.to_device(z_hat.astype(np.float32)), 1e-9)
        ys = self.sc.decompress_cubes(list(y_strings), locs, scales, y_min, y_max)
        xs = c.synthesis(ys.reshape(len(y_strings), 16, 16, 16, 16))
        ks = np.array([int(rho * np.array(n)) for n in nums], np.int32)
        mask, _, _ = c.topk(xs, c.to_device(ks))
        return self.runtime.to_host(mask, "mask").copy()


def _dist():
    import torch.distributed as dist
    return dist


def compress_sharded(cubes_local: np.ndarray, local: LocalCodec, group=None) -> Optional[dict]:
    """Every rank passes ITS slice (``shard_slices`` order).  Rank 0 returns the stream of the whole cloud
    {'y_strings', 'y_min', 'y_max', 'y_shape', 'z_string', 'z_min', 'z_max', 'z_shape'}; other ranks None."""
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    part = local.encode_local(cubes_local)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0, group=group)
    if rank != 0:
        return None
    y_strings = [s for p in parts for s in p["y_strings"]]
    z_hat = np.concatenate([p["z_hat"] for p in parts])
    z_string, z_min, z_max = local.encode_z(z_hat)
    return {"y_strings": y_strings, "y_min": np.concatenate([p["y_min"] for p in parts]),
            "y_max": np.concatenate([p["y_max"] for p in parts]), "y_shape": np.array([1, 16, 16, 16, 16], np.int64),
            "z_string": z_string, "z_min": z_min, "z_max": z_max, "z_shape": np.array(z_hat.shape, np.int32)}


def decompress_sharded(stream: Optional[dict], nums: Optional[np.ndarray], rho: float, local: LocalCodec, group=None) -> Optional[np.ndarray]:
    """Rank 0 passes the stream (+ the per-cube point counts); every rank decodes its slice; rank 0 returns the
    uint8 occupancy masks of all cubes in order."""
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pieces = None
    if rank == 0:
        B = len(stream["y_strings"])
        z_hat = local.decode_z(stream["z_string"], stream["z_min"], stream["z_max"], stream["z_shape"])
        pieces = []
        for a, b in shard_slices(B, world):
            pieces.append({"y_strings": stream["y_strings"][a:b], "y_min": stream["y_min"][a:b], "y_max": stream["y_max"][a:b],
                           "z_hat": z_hat[a:b], "nums": np.asarray(nums)[a:b]})
    mine = [None]
    dist.scatter_object_list(mine, pieces, src=0, group=group)
    m = mine[0]
    mask = local.decode_local(m["y_strings"], m["y_min"], m["y_max"], m["z_hat"], m["nums"], rho)
    masks = [None] * world if rank == 0 else None
    dist.gather_object(mask, masks, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate(masks)
