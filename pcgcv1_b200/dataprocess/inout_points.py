"""Drop-in for ``dataprocess/inout_points.py``: ``select_voxels`` (top-k occupancy classification, :147-179),
``voxels2points`` (:134-143), ``points2voxels`` (:116-132) on the codec path, and -- SURVEY.md section 8(f) rank 1 -- the
callers either side of it: ``load_ply_data`` (:8-28), ``write_ply_data`` (:30-46), ``load_points`` (:50-90) and
``save_points`` (:92-112) with the reference's signatures, return values, ordering and filtering quirks.

``select_voxels`` runs the radix-select kernel of libpcgc_b200.so; the PLY parser / writer and the cube partition are the
library's multithreaded C++ (``csrc/pointio.cpp``) in place of per-line / per-point Python; the ``*_packed`` and
``*_device`` variants keep points as (array, offsets) and cubes / masks on the GPU so that 6 bytes per point cross PCIe
instead of 256 KiB - 2 MiB per cube."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .. import _lib, runtime


# ---------------------------------------------------------------- plyfile <--> points
def load_ply_data(filename):
    """ASCII .ply -> int32 [n,3]: every line whose first three space-separated tokens are floats (inout_points.py:8-28)."""
    with open(filename, "rb") as f:
        data = f.read()
    L = _lib.lib()
    cap = data.count(b"\n") + 1
    xyz = np.empty((cap, 3), np.int32)
    n = C.c_int64()
    rc = L.pcgc_ply_parse(data, len(data), xyz.ctypes.data, cap, C.byref(n), 0)
    if rc == _lib.ERR_CORRUPT:
        raise IndexError("list index out of range")           # a numeric line with fewer than three tokens (:18)
    _lib.check(rc)
    if n.value == 0:
        return np.array([]).astype(np.int32)                  # what np.array([]) gives the reference
    return xyz[:n.value].copy() if n.value < cap // 2 else xyz[:n.value]


def write_ply_data(filename, points):
    """[n,3] -> ASCII .ply, byte for byte the reference's file (inout_points.py:30-46)."""
    points = np.asarray(points)
    if os.path.exists(filename):
        os.remove(filename)
    if points.ndim == 2 and points.shape[1] >= 3 and points.dtype.kind in "iu" and \
            (points.size == 0 or (points.min() >= -2**31 and points.max() < 2**31)):
        xyz = np.ascontiguousarray(points[:, :3], dtype=np.int32)
        cap = 160 + 36 * len(xyz)
        out = np.empty(cap, np.uint8)
        ln = C.c_int64()
        _lib.check(_lib.lib().pcgc_ply_format(xyz.ctypes.data, len(xyz), out.ctypes.data, cap, C.byref(ln), 0))
        with open(filename, "wb") as f:
            f.write(memoryview(out[:ln.value]))
        return
    # float coordinates (process.py:27-31,74-77): Python's str() of every scalar, like the reference
    with open(filename, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex " + str(points.shape[0]) + "\n")
        f.write("property float x\nproperty float y\nproperty float z\nend_header\n")
        f.write("".join([str(p[0]) + " " + str(p[1]) + " " + str(p[2]) + "\n" for p in points]))


# ---------------------------------------------------------------- plyfile <--> partitioned points
def partition_points(point_cloud, cube_size=64, min_num=20):
    """int [n,3] -> (local int16 [m,3] grouped per kept cube in sorted cube order, offsets int64 [B+1],
    cube_positions [B,3] in first-appearance order (what the reference returns), cube_positions_sorted [B,3])."""
    xyz = np.ascontiguousarray(np.asarray(point_cloud).reshape(-1, 3), dtype=np.int32)
    n = len(xyz)
    seen = np.empty((max(n, 1), 3), np.int64)
    srt = np.empty((max(n, 1), 3), np.int64)
    counts = np.empty(max(n, 1), np.int64)
    nc, npts = C.c_int64(), C.c_int64()
    cap = max(n, 1)
    while True:
        local = np.empty((cap, 3), np.int16)
        rc = _lib.lib().pcgc_partition_points(xyz.ctypes.data, n, int(cube_size), int(min_num), local.ctypes.data, cap, seen.ctypes.data,
                                               srt.ctypes.data, counts.ctypes.data, C.byref(nc), C.byref(npts))
        if rc != _lib.ERR_OVERFLOW:
            break
        cap = npts.value                                          # a repeated cube (negative coordinates, see csrc/pointio.cpp)
    if rc == _lib.ERR_BAD_RANGE:
        raise KeyError("cube position order cannot be decoded (negative cube coordinates)")
    _lib.check(rc)
    B = nc.value
    offsets = np.zeros(B + 1, np.int64)
    np.cumsum(counts[:B], out=offsets[1:])
    return local[:npts.value].copy(), offsets, seen[:B].copy(), srt[:B].copy()


def load_points(filename, cube_size=64, min_num=20):
    """.ply -> (set_points: list of int16 [n_i,3] in sorted cube order, cube_positions [B,3] in first-appearance order)
    exactly as inout_points.py:50-90 (including the cube_positions order the reference returns)."""
    local, offsets, seen, _ = partition_points(load_ply_data(filename), cube_size, min_num)
    if len(seen) == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")   # cube_positions.max() (:78)
    set_points = [local[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]
    set_points = [p.reshape(3) if len(p) == 1 else p for p in set_points]    # a single point stays 1-D in the reference (:66)
    return set_points, seen


def load_points_packed(filename, cube_size=64, min_num=20):
    """As load_points but packed: (local int16 [m,3], offsets int64 [B+1], cube_positions [B,3])."""
    local, offsets, seen, _ = partition_points(load_ply_data(filename), cube_size, min_num)
    if len(seen) == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    return local, offsets, seen


def _ordered_positions(cube_positions):
    cube_positions = np.asarray(cube_positions)
    step = cube_positions.max() + 1
    n = cube_positions[:, 0:1] + cube_positions[:, 1:2] * step + cube_positions[:, 2:3] * step * step
    n = np.sort(n, axis=0)
    return np.concatenate((n % step, (n // step) % step, n // step // step), -1)


def save_points(set_points, cube_positions, filename, cube_size=64):
    """Combine the per-cube points (in sorted cube order) with their cube positions and write the .ply (:92-112)."""
    ordered = _ordered_positions(cube_positions)
    parts = [np.asarray(v) + np.array(k) * cube_size for k, v in zip(ordered, set_points)]
    write_ply_data(filename, np.concatenate(parts).astype("int"))


def save_points_packed(points, counts, cube_positions, filename, cube_size=64):
    """save_points for points packed as one [n,3] array + per-cube counts."""
    ordered = _ordered_positions(cube_positions)
    counts = np.asarray(counts).reshape(-1)
    if len(counts) > len(ordered):
        counts = counts[:len(ordered)]                           # zip() truncation of the reference
    m = int(counts.sum())
    pc = np.asarray(points[:m]).astype(np.int64) + np.repeat(ordered[:len(counts)].astype(np.int64) * cube_size, counts, axis=0)
    write_ply_data(filename, pc.astype("int"))


def select_voxels(vols, points_nums, offset_ratio=1.0, fixed_thres=None, codec=None, dtype="float32"):
    """vols [B,S,S,S,1] float32 logits (NumPy, torch or a DeviceResult), points_nums [B]
    -> float32 mask [B,S,S,S,1] of the top ``int(offset_ratio*points_nums[b])`` voxels per cube, ties
    included (``>=``), as NumPy like the reference.  ``dtype="uint8"`` skips the float32 widening that the
    reference's own consumer undoes again (voxels2points casts to uint8, inout_points.py:137)."""
    c = codec or runtime.get_codec("voxception", "")
    if isinstance(vols, runtime.PendingDeviceResult) and not vols.finalized and fixed_thres is None and len(vols) > 0:
        return _select_voxels_pipelined(vols._codec, vols, points_nums, offset_ratio, dtype)
    v = c.to_device(vols, torch.float32)
    B = v.shape[0]
    if B == 0:                                                           # the reference's loop over zero cubes: an empty mask
        return np.zeros(tuple(v.shape), np.dtype(dtype))
    if fixed_thres is None:
        pn = np.asarray(runtime.unwrap(points_nums)).reshape(-1)
        ks = np.array([int(offset_ratio * np.array(pn[i])) for i in range(B)], np.int32)
        if (ks > v[0].numel()).any():
            raise IndexError("select_voxels: k exceeds the number of voxels (get_adaptive_thres would raise IndexError)")
        mask, _, _ = c.topk(v, c.to_device(ks))
    else:
        mask, _ = c.threshold(v, float(fixed_thres))
    m = runtime.to_host(mask, "mask")
    return m.astype(dtype) if np.dtype(dtype) != m.dtype else runtime.host_copy(m)


def _select_voxels_pipelined(c, pend, points_nums, offset_ratio, dtype):
    """select_voxels on a result that is still being produced (decompress_hyper's PendingDeviceResult): top-k and the mask's
    device -> host copy of cubes [a, b) start when THAT part's synthesis has finished, beside the synthesis of the later parts;
    the host copies a part out of the pinned staging buffer while the GPU works on the next one.  Same kernel per cube as the
    plain path, so the masks are identical."""
    full = pend.raw
    B = full.shape[0]
    V = full[0].numel()
    pn = np.asarray(runtime.unwrap(points_nums)).reshape(-1)
    ks = np.array([int(offset_ratio * np.array(pn[i])) for i in range(B)], np.int32)
    if (ks > V).any():
        raise IndexError("select_voxels: k exceeds the number of voxels (get_adaptive_thres would raise IndexError)")
    dev = full.device
    side, cs = c.coder_stream(7), runtime.copy_stream(dev)
    stage = runtime.pinned_buffer("mask", B * V)[:B * V].view(B, V)
    out = np.empty(tuple(full.shape), np.uint8)
    out2 = out.reshape(B, V)
    runtime.COUNTERS["d2h_bytes"] += B * V
    jobs = []
    c.deferred_checks(True)
    try:
        with torch.cuda.stream(side):
            ksd = c.to_device(ks)
        for a, b, ev in pend.parts:
            with torch.cuda.stream(side):
                side.wait_event(ev)
                m, _, _ = c.topk(full[a:b], ksd[a:b])
                ready = torch.cuda.Event()
                ready.record(side)
            m.record_stream(cs)
            with torch.cuda.stream(cs):
                cs.wait_event(ready)
                stage[a:b].copy_(m.view(b - a, V), non_blocking=True)
                done = torch.cuda.Event()
                done.record(cs)
            jobs.append((a, b, done))
        for a, b, done in jobs:
            done.synchronize()
            runtime.host_copy_into(out2[a:b], stage[a:b].numpy())
    finally:
        c.deferred_checks(False)
    pend.finalize()                                      # whole section done; raises if a kernel flagged an error
    return out if np.dtype(dtype) == np.uint8 else out.astype(dtype)


def select_voxels_device(codec, logits: torch.Tensor, ks: torch.Tensor):
    """Device-resident form: -> (mask uint8, thres, count) torch tensors."""
    return codec.topk(logits, ks)


def points2voxels_device(local, offsets, cube_size=64, codec=None):
    """Packed points -> DeviceResult of the uint8 occupancy cubes [B,S,S,S,1] on the GPU."""
    c = codec or runtime.get_codec("voxception", "")
    return runtime.DeviceResult(c.voxelize(local, offsets, int(cube_size)))


def voxels2points_device(mask, codec=None, cap=None):
    """Device mask (uint8 [B,S,S,S,1]) -> (points int16 [n,3], counts int32 [B]) in voxels2points' order."""
    c = codec or runtime.get_codec("voxception", "")
    return c.extract_points(runtime.unwrap(mask), cap)


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
