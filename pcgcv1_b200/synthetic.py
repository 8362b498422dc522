"""Seeded synthetic workloads (BASELINE.json configs 1-4) -- host side, NumPy only.

There is no network and no test cloud in the reference tree, so the named workloads are
generated: a dense "vox10" surface (1024^3 grid, ~800k points, ~200 cubes of 64^3, like
longdress_vox10 in demo.ipynb cell 9) and a sparse "vox12" scan (4096^3 grid, thousands of
lightly filled cubes).  Partitioning follows ``dataprocess/inout_points.py:50-90`` (cube index
= point // cube_size, cubes with < min_num points dropped, cubes ordered by
``x + y*step + z*step^2``) but vectorised; voxelisation follows ``points2voxels`` (``:116-132``)
and emits uint8 occupancy instead of float64.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def cloud_vox10(seed: int = 0, target_points: int = 800_000) -> np.ndarray:
    """Closed wavy surface in a 1024^3 grid, unique integer points, ~target_points."""
    rng = np.random.default_rng(seed)
    n_theta, n_phi = 2000, 4000
    theta = (np.arange(n_theta) + 0.5) * (np.pi / n_theta)
    phi = (np.arange(n_phi) + 0.5) * (2 * np.pi / n_phi)
    t, p = np.meshgrid(theta, phi, indexing="ij")
    ph = rng.uniform(0, 2 * np.pi, 3)
    r0 = 172.0 * np.sqrt(target_points / 800_000.0)
    r = r0 * (1.0 + 0.08 * np.sin(5 * t + ph[0]) * np.cos(4 * p + ph[1]) + 0.04 * np.sin(11 * p + ph[2]))
    x = 512 + r * np.sin(t) * np.cos(p)
    y = 512 + r * np.sin(t) * np.sin(p) * 1.15
    z = 512 + r * np.cos(t) * 1.3
    pts = np.stack([x, y, z], -1).reshape(-1, 3)
    pts = np.clip(np.rint(pts), 0, 1023).astype(np.int64)
    return _unique_points(pts, 1024)


def cloud_vox12(seed: int = 0, n_sheets: int = 14, keep: float = 0.07) -> np.ndarray:
    """Sparse facade-like sheets with random drop-out in a 4096^3 grid."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_sheets):
        o = rng.uniform(300, 3700, 3)
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        v = rng.normal(size=3); v -= u * (u @ v); v /= np.linalg.norm(v)
        w = np.cross(u, v)
        ext = rng.uniform(900, 1800, 2)
        n = int(ext[0] * ext[1] * keep)
        a = rng.uniform(-ext[0] / 2, ext[0] / 2, n)
        b = rng.uniform(-ext[1] / 2, ext[1] / 2, n)
        bump = 6.0 * np.sin(a / 90.0) * np.cos(b / 70.0)
        pts = o + a[:, None] * u + b[:, None] * v + bump[:, None] * w
        out.append(pts)
    pts = np.rint(np.concatenate(out)).astype(np.int64)
    pts = pts[np.all((pts >= 0) & (pts < 4096), axis=1)]
    return _unique_points(pts, 4096)


def _unique_points(pts: np.ndarray, res: int) -> np.ndarray:
    key = np.unique((pts[:, 0] * res + pts[:, 1]) * res + pts[:, 2])
    return np.stack([key // (res * res), (key // res) % res, key % res], -1).astype(np.int32)


def partition(points: np.ndarray, cube_size: int = 64, min_num: int = 64
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """points [n,3] int -> (cubes uint8 [B,S,S,S,1], cube_positions int32 [B,3], points_numbers uint16 [B]).

    Same result as ``load_points`` + ``points2voxels`` + the uint16 count of ``process.py:45``,
    cubes in the reference's sorted order."""
    points = np.asarray(points).astype(np.int64)
    idx = points // cube_size
    local = points % cube_size
    step = int(idx.max()) + 1
    key = idx[:, 0] + idx[:, 1] * step + idx[:, 2] * step * step
    uniq, inv, counts = np.unique(key, return_inverse=True, return_counts=True)
    keep = counts >= min_num
    new_id = np.cumsum(keep) - 1
    sel = keep[inv]
    b = new_id[inv[sel]]
    lp = local[sel]
    nb = int(keep.sum())
    cubes = np.zeros((nb, cube_size, cube_size, cube_size, 1), np.uint8)
    cubes[b, lp[:, 0], lp[:, 1], lp[:, 2], 0] = 1
    k = uniq[keep]
    # NOTE: load_points re-derives x,y,z with a step computed AFTER filtering; equal unless the
    # largest index only appears in dropped cubes.  We keep the pre-filter step for both.
    pos = np.stack([k % step, (k // step) % step, k // step // step], -1).astype(np.int32)
    nums = cubes.reshape(nb, -1).sum(1).astype(np.uint16)
    return cubes, pos, nums


def workload(name: str, seed: int = 0, max_cubes: int | None = None):
    """name in {'vox10', 'vox12'} -> (cubes, cube_positions, points_numbers)."""
    if name == "vox10":
        pts = cloud_vox10(seed)
    elif name == "vox12":
        pts = cloud_vox12(seed)
    else:
        raise ValueError(name)
    cubes, pos, nums = partition(pts, 64, 64)
    if max_cubes is not None:
        cubes, pos, nums = cubes[:max_cubes], pos[:max_cubes], nums[:max_cubes]
    return cubes, pos, nums


def surface_cubes(n: int, seed: int = 0, cube_size: int = 64) -> Tuple[np.ndarray, np.ndarray]:
    """n independent random smooth-surface cubes (config 4 batch sweep, small tests)."""
    rng = np.random.default_rng(seed)
    g = np.arange(cube_size)
    xx, yy = np.meshgrid(g, g, indexing="ij")
    cubes = np.zeros((n, cube_size, cube_size, cube_size, 1), np.uint8)
    for i in range(n):
        a = rng.uniform(-0.6, 0.6, 2)
        f = rng.uniform(0.05, 0.25, 2)
        ph = rng.uniform(0, 6.28, 2)
        h = cube_size / 2 + a[0] * (xx - 32) + a[1] * (yy - 32) + 5 * np.sin(f[0] * xx + ph[0]) * np.cos(f[1] * yy + ph[1])
        hz = np.rint(h).astype(np.int64)
        ok = (hz >= 0) & (hz < cube_size)
        perm = rng.permutation(3)
        coords = [xx[ok], yy[ok], hz[ok]]
        cubes[i, coords[perm[0]], coords[perm[1]], coords[perm[2]], 0] = 1
        # second thin layer to vary the density
        if rng.random() < 0.5:
            hz2 = np.clip(hz + 1, 0, cube_size - 1)
            coords = [xx[ok], yy[ok], hz2[ok]]
            cubes[i, coords[perm[0]], coords[perm[1]], coords[perm[2]], 0] = 1
    nums = cubes.reshape(n, -1).sum(1).astype(np.uint16)
    return cubes, nums
