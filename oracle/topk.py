"""Oracle restatement of the decoder's top-k occupancy classification (test infrastructure only).

Follows ``dataprocess/inout_points.py:134-179`` statement by statement (NumPy, like the
reference).  Pinned: ``tests/golden/make_golden.py`` imports the reference's own
``inout_points`` module (pure NumPy, importable offline) and the golden masks are committed.
"""
from __future__ import annotations

import numpy as np


def get_adaptive_thres(vol: np.ndarray, num: int, init_thres: float = -2.0):
    """inout_points.py:170-179.  Quirks kept: ``num == 0`` indexes ``values[-0] = values[0]``
    (the minimum); ``num > size`` raises IndexError."""
    values = vol[vol > init_thres]
    if values.shape[0] < num:
        values = np.reshape(vol, [-1])
    values = np.sort(values)
    return values[-num]


def select_voxels(vols: np.ndarray, points_nums, offset_ratio: float = 1.0, fixed_thres=None) -> np.ndarray:
    """inout_points.py:147-168: mask = (vol >= k-th largest), k = int(rho * n_points)."""
    masks = []
    for idx, vol in enumerate(vols):
        if fixed_thres is None:
            num = int(offset_ratio * np.array(points_nums[idx]))
            thres = get_adaptive_thres(vol, num)
        else:
            thres = fixed_thres
        masks.append(np.greater_equal(vol, thres).astype("float32"))
    return np.stack(masks)


def voxels2points(voxels: np.ndarray):
    """inout_points.py:134-143: lexicographic (d,h,w) coordinates of the set voxels."""
    voxels = np.squeeze(np.uint8(voxels))
    if voxels.ndim == 3:
        voxels = voxels[None]
    return [np.array(np.where(vol > 0)).transpose((1, 0)) for vol in voxels]


def points2voxels(set_points, cube_size: int) -> np.ndarray:
    """inout_points.py:116-132."""
    voxels = []
    for points in set_points:
        points = points.astype("int")
        vol = np.zeros((cube_size, cube_size, cube_size))
        vol[points[:, 0], points[:, 1], points[:, 2]] = 1.0
        voxels.append(np.expand_dims(vol, -1))
    return np.array(voxels)
