"""Pre- and post-processing around the codec with the reference's entry points: ``preprocess(input_file, scale, cube_size,
min_num)`` -> ``(cubes, cube_positions, points_numbers)`` and ``postprocess(output_file, cubes, points_numbers, cube_positions,
scale, cube_size, rho, fixed_thres=None)`` (reference ``process.py:16-52`` and ``:54-82``).

What happens underneath: the cube partition is the library's C++ (6.3 s -> 0.05 s on the vox10
cloud); the occupancy cubes are built ON the GPU from the packed points and handed to the codec as a device handle (the
reference materialises a float64 [B,64,64,64,1] host array: 2 MiB per cube); on the way back the top-k mask never leaves the
GPU -- an ordered compaction returns the coordinate list, which the C++ writer turns into the .ply."""
from __future__ import annotations

import os
import time
import uuid

import numpy as np

from . import runtime
from .dataprocess import inout_points as iop


class _Stage:
    """``with _Stage("Partition"):`` prints the elapsed seconds of a stage like the reference's progress lines."""

    def __init__(self, label):
        self.label = label

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            print("%s: %.4fs" % (self.label, time.time() - self.t0))
        return False


def _scratch(tag):
    return "%s_%s.ply" % (tag, uuid.uuid4().hex[:12])


def _rescaled_copy(src, factor, unique):
    """Writes ``round(points * factor)`` (de-duplicated for down-scaling) to a scratch .ply and returns its name."""
    pc = iop.load_ply_data(src).astype("float32") * factor
    if unique:
        pc = np.unique(np.round(pc), axis=0)                    # remove the points that collapse onto each other
    dst = _scratch("rescaled")
    iop.write_ply_data(dst, pc)
    return dst


def preprocess(input_file, scale, cube_size, min_num, codec=None):
    """Optional down-scaling, cube partition, voxelisation, point counts.  ``cubes`` is a device handle (uint8 [B,S,S,S,1];
    ``.numpy()`` / ``.shape`` like the array the reference returns, same values); ``cube_positions`` comes back in
    first-appearance order and ``points_numbers`` as uint16, both as in the reference."""
    c = codec or runtime.get_codec("voxception", "")
    source = input_file
    with _Stage("Scaling"):
        if scale != 1:
            source = _rescaled_copy(input_file, scale, unique=True)
    try:
        with _Stage("Partition"):
            local, offsets, cube_positions = iop.load_points_packed(source, cube_size, min_num)
    finally:
        if source != input_file:
            os.remove(source)
    with _Stage("Voxelization"):
        cubes = iop.points2voxels_device(local, offsets, cube_size, codec=c)
        points_numbers = c.count_voxels(cubes.tensor).astype(np.uint16)
    print("cubes %s; points per cube: total %d, mean %d, max %d, min %d" % (
        cubes.shape, points_numbers.sum(), round(points_numbers.mean()), points_numbers.max(), points_numbers.min()))
    return cubes, cube_positions, points_numbers


def postprocess(output_file, cubes, points_numbers, cube_positions, scale, cube_size, rho, fixed_thres=None, codec=None):
    """Occupancy classification (the ``rho * points_numbers[i]`` most likely voxels of cube i, or a fixed threshold), point
    extraction, re-assembly of the cubes and the .ply (scaled back up when ``scale != 1``)."""
    import torch
    c = codec or runtime.get_codec("voxception", "")
    with _Stage("Classify and extract points"):
        logits = c.to_device(cubes, torch.float32)
        if fixed_thres is not None:
            mask, _ = c.threshold(logits, float(fixed_thres))
        else:
            counts = np.asarray(runtime.unwrap(points_numbers)).reshape(-1)
            ks = np.array([int(rho * np.array(counts[i])) for i in range(logits.shape[0])], np.int32)
            if (ks > logits[0].numel()).any():
                raise IndexError("select_voxels: k exceeds the number of voxels (get_adaptive_thres would raise IndexError)")
            mask, _, _ = c.topk(logits, c.to_device(ks))
        points, per_cube = iop.voxels2points_device(mask, codec=c)
    with _Stage("Write point cloud to %s" % output_file):
        if scale == 1:
            iop.save_points_packed(points, per_cube, cube_positions, output_file, cube_size)
        else:
            tmp = _scratch("reconstruction")
            try:
                iop.save_points_packed(points, per_cube, cube_positions, tmp, cube_size)
                iop.write_ply_data(output_file, iop.load_ply_data(tmp).astype("float32") * float(1 / scale))
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
