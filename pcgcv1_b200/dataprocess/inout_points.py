"""Drop-in for the hot-path part of ``dataprocess/inout_points.py``: ``select_voxels`` (top-k occupancy
classification, :147-179), ``voxels2points`` (:134-143) and ``points2voxels`` (:116-132).

``select_voxels`` runs the radix-select kernel of libpcgc_b200.so; the PLY reader/writer and the cube
partitioner of the same reference file are host I/O outside this path."""
from __future__ import annotations

import numpy as np
import torch

from .. import runtime


def select_voxels(vols, points_nums, offset_ratio=1.0, fixed_thres=None, codec=None, dtype="float32"):
    """vols [B,S,S,S,1] float32 logits (NumPy, torch or a DeviceResult), points_nums [B]
    -> float32 mask [B,S,S,S,1] of the top ``int(offset_ratio*points_nums[b])`` voxels per cube, ties
    included (``>=``), as NumPy like the reference.  ``dtype="uint8"`` skips the float32 widening that the
    reference's own consumer undoes again (voxels2points casts to uint8, inout_points.py:137)."""
    c = codec or runtime.get_codec("voxception", "")
    v = c.to_device(vols, torch.float32)
    B = v.shape[0]
    if fixed_thres is None:
        pn = np.asarray(runtime.unwrap(points_nums)).reshape(-1)
        ks = np.array([int(offset_ratio * np.array(pn[i])) for i in range(B)], np.int32)
        if (ks > v[0].numel()).any():
            raise IndexError("select_voxels: k exceeds the number of voxels (get_adaptive_thres would raise IndexError)")
        mask, _, _ = c.topk(v, c.to_device(ks))
    else:
        mask, _ = c.threshold(v, float(fixed_thres))
    m = runtime.to_host(mask, "mask")
    return m.astype(dtype) if np.dtype(dtype) != m.dtype else m.copy()


def select_voxels_device(codec, logits: torch.Tensor, ks: torch.Tensor):
    """Device-resident form: -> (mask uint8, thres, count) torch tensors."""
    return codec.topk(logits, ks)


def voxels2points(voxels):
    """[B,S,S,S,1] 0/1 -> list of [n_i,3] integer coordinates in lexicographic (d,h,w) order."""
    voxels = np.uint8(np.asarray(runtime.unwrap(voxels)))
    if voxels.ndim == 5:
        voxels = voxels[..., 0]
    elif voxels.ndim == 3:
        voxels = voxels[None]
    return [np.array(np.where(vol > 0)).transpose((1, 0)) for vol in voxels]


def points2voxels(set_points, cube_size):
    """list of [n_i,3] -> uint8 occupancy [B,S,S,S,1] (the reference builds float64; values equal)."""
    out = np.zeros((len(set_points), cube_size, cube_size, cube_size, 1), np.uint8)
    for i, points in enumerate(set_points):
        p = np.asarray(points).astype("int")
        out[i, p[:, 0], p[:, 1], p[:, 2], 0] = 1
    return out
