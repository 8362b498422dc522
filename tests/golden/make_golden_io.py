"""Golden vectors for the I/O rows: made by EXECUTING THE REFERENCE'S OWN
``dataprocess/inout_points.py`` (pure NumPy / Python) and ``dataprocess/inout_bitstream.py`` on seeded inputs.

    python tests/golden/make_golden_io.py       # needs /root/reference (read-only), writes golden_io.npz here

The only patch: ``gpcc_encode`` / ``gpcc_decode`` (a subprocess call to the prebuilt ``myutils/tmc3`` binary) are replaced by
a file copy, so the golden covers the container layout of the other four files and the temp ``_cubepos.ply``.
/root/reference does not exist on the GPU box: tests only read the .npz written here."""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCGC_REFERENCE", "/root/reference")


def make_ply_text(rng) -> str:
    """A small ASCII .ply with the things the reference's parser meets: header words, a numeric-looking comment, extra
    columns, floats, exponents, a blank line, a double space (skipped), duplicates, one negative point, no final newline."""
    pts = rng.integers(0, 200, size=(6000, 3))
    pts = np.concatenate([pts, pts[:40]])
    lines = ["ply", "format ascii 1.0", "comment 1 2 x", "element vertex %d" % len(pts), "property float x", "property float y",
             "property float z", "property uchar red", "end_header"]
    for i, p in enumerate(pts):
        if i % 3 == 0:
            lines.append("%d %d %d 255" % tuple(p))
        elif i % 3 == 1:
            lines.append("%.1f %d %de0 7 8" % tuple(p))
        else:
            lines.append("%d %d %d" % tuple(p))
    lines += ["", "3.9 -2.7 1e2", "1  2 3", "7 8 9"]
    return "\n".join(lines)


def main():
    sys.path.insert(0, REF)
    iop = importlib.import_module("dataprocess.inout_points")
    ibs = importlib.import_module("dataprocess.inout_bitstream")
    ibs.gpcc_encode = lambda ply, out, show=False: shutil.copyfile(ply, out)
    ibs.gpcc_decode = lambda binf, ply, show=False: shutil.copyfile(binf, ply)
    rng = np.random.default_rng(20260101)
    out = {}
    tmp = tempfile.mkdtemp(prefix="pcgc_golden_io_")
    text = make_ply_text(rng)
    ply = os.path.join(tmp, "cloud.ply")
    with open(ply, "w") as f:
        f.write(text)
    out["ply_text"] = np.frombuffer(text.encode(), np.uint8)
    out["load_ply_data"] = iop.load_ply_data(ply)
    for mn in (1, 3, 20, 64):
        sp, cp = iop.load_points(ply, 64, mn)
        out["lp%d_counts" % mn] = np.array([1 if p.ndim == 1 else len(p) for p in sp], np.int64)
        out["lp%d_ndim" % mn] = np.array([p.ndim for p in sp], np.int64)
        out["lp%d_points" % mn] = np.concatenate([p.reshape(-1, 3) for p in sp]).astype(np.int16)
        out["lp%d_cube_positions" % mn] = cp
    sp, cp = iop.load_points(ply, 64, 20)
    rec = os.path.join(tmp, "rec.ply")
    iop.save_points(sp, cp, rec, 64)
    out["save_points_bytes"] = np.frombuffer(open(rec, "rb").read(), np.uint8)
    vox = iop.points2voxels(sp, 64)
    out["points_numbers"] = np.sum(vox, axis=(1, 2, 3, 4)).astype(np.uint16)
    fl = os.path.join(tmp, "float.ply")
    iop.write_ply_data(fl, out["load_ply_data"][:64].astype("float32") * float(1 / 3))
    out["write_float_bytes"] = np.frombuffer(open(fl, "rb").read(), np.uint8)

    # ---- bitstream container
    B = len(sp)
    lens = rng.integers(1, 600, size=B)
    lens[0], lens[1] = 255, 256
    y_strings = [bytes(rng.integers(1, 256, size=int(l), dtype=np.uint8)) for l in lens]     # no trailing NULs (np.array(bytes) strips them)
    y_min = -rng.integers(0, 16, size=B).astype(np.int32)
    y_max = rng.integers(0, 16, size=B).astype(np.int32)
    z_string = bytes(rng.integers(1, 256, size=777, dtype=np.uint8))
    pn = out["points_numbers"]
    y_shape = np.array([1, 16, 16, 16, 16])
    z_shape = np.array([B, 8, 8, 8, 8])
    root = os.path.join(tmp, "compressed")
    ibs.write_binary_files_hyper("g", np.array(y_strings, dtype=object), z_string, pn, cp, y_min, y_max, y_shape, np.int32(-9), np.int32(11),
                                 z_shape, rootdir=root)
    for ext in (".strings", ".strings_head", ".strings_hyper", ".pointnums", "_cubepos.ply"):
        out["hyper" + ext] = np.frombuffer(open(os.path.join(root, "g" + ext), "rb").read(), np.uint8)
    # (the reference's READER does not run under NumPy 2 -- it builds an int32 array from a mix of shape-(1,) arrays and
    #  ints, inout_bitstream.py:168-174 -- so the reader is pinned through the files its WRITER produced above)
    out["hyper_in_y_lens"] = lens.astype(np.int64)
    out["hyper_in_y_concat"] = np.frombuffer(b"".join(y_strings), np.uint8)
    out["hyper_in_y_min"], out["hyper_in_y_max"] = y_min, y_max
    out["hyper_in_z"] = np.frombuffer(z_string, np.uint8)
    out["cube_positions"] = cp
    f_string = bytes(rng.integers(1, 256, size=3210, dtype=np.uint8))
    ibs.write_binary_files_factorized("f", f_string, pn, cp, np.int32(-20), np.int32(17), np.array([B, 8, 8, 8, 32]), rootdir=root)
    for ext in (".strings", ".pointnums", "_cubepos.ply"):
        out["fact" + ext] = np.frombuffer(open(os.path.join(root, "f" + ext), "rb").read(), np.uint8)
    out["fact_in_string"] = np.frombuffer(f_string, np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_io.npz"), **out)
    shutil.rmtree(tmp)
    print("wrote golden_io.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
