"""On-disk container of the codec, interface-compatible with the reference's ``dataprocess/inout_bitstream.py``
(``write_binary_files_factorized`` :10-44, ``read_binary_files_factorized`` :46-70, ``write_binary_files_hyper`` :75-141,
``read_binary_files_hyper`` :144-198; layout in SURVEY.md Appendix B).  The CONTAINER (headers, length fields, side files) written by either side reads on the other; the range-coded payloads are
byte-identical to a literal restatement of ``tensorflow.contrib.coder``'s RangeEncoder (oracle/coder.py, tests/test_host_coder.py); that
restatement against a TF binary is unpinned (no TF wheel and no golden stream in the reference), so payload interchange with a TF 1.13 run
is untested.

Files next to ``<rootdir>/<filename>``:

=================  ==================================================================================================
``.strings``       factorized: ``int16 shape[5] | int8 min_v | int8 max_v | string``;  hyper: the y strings back to back
``.strings_head``  hyper: ``int16 n | uint8 (max_v*16 - min_v) x n | length x n | int16 y_shape[5]``; a length is one ``uint8`` if
                   1..255, else a ``0`` byte followed by 16 bits
``.strings_hyper`` hyper: ``int16 z_shape[5] | int8 z_min_v | int8 z_max_v | z string``
``.pointnums``     ``uint16`` per cube
``.cubepos``       cube positions through ``myutils.gpcc_wrapper`` (temporary ``_cubepos.ply`` beside it)
=================  ==================================================================================================

Where the reference would silently wrap a value that does not fit its field, this module raises ``ValueError``: more than
32 767 cubes, y symbol ranges outside [-15, 0] / [0, 15], global ranges outside int8, cube coordinates above 255, strings
of 65 536 bytes or more.  Two deliberate extensions, both invisible to streams the reference can write: the 16-bit length
after a ``0`` byte is read unsigned (random-init weights code about 42 KB per cube; the reference's int16 overflows at
32 768), and an EMPTY string is stored in that escape form (the reference writes a bare ``0`` its own reader cannot parse).
"""
from __future__ import annotations

import os
from collections import namedtuple

import numpy as np

from ..myutils.gpcc_wrapper import gpcc_decode, gpcc_encode
from .inout_points import load_ply_data, write_ply_data

_Paths = namedtuple("_Paths", "strings head hyper pointnums cubepos cubepos_ply")
_I16, _I8, _U8, _U16 = np.dtype("<i2"), np.dtype("i1"), np.dtype("u1"), np.dtype("<u2")


def _paths(filename, rootdir) -> _Paths:
    stem = os.path.join(rootdir, filename)
    return _Paths(stem + ".strings", stem + ".strings_head", stem + ".strings_hyper", stem + ".pointnums", stem + ".cubepos",
                  stem + "_cubepos.ply")


def _blob(s) -> bytes:
    """bytes from bytes / np.bytes_ / 0-d or object arrays / anything with ``.numpy()``."""
    s = s.numpy() if hasattr(s, "numpy") else s
    if isinstance(s, np.ndarray):
        s = s.item() if (s.dtype == object or s.ndim == 0) else s.tobytes()
    return bytes(s)


def _int8_pair(lo, hi, what) -> bytes:
    lo, hi = int(np.asarray(lo)), int(np.asarray(hi))
    if not (-128 <= lo <= 127 and -128 <= hi <= 127):
        raise ValueError("%s symbol range [%d, %d] does not fit the int8 header" % (what, lo, hi))
    return np.array((lo, hi), dtype=_I8).tobytes()


def _shape5(shape) -> bytes:
    return np.asarray(shape).astype(_I16).reshape(-1).tobytes()


def _put(path, *chunks):
    with open(path, "wb") as f:
        for c in chunks:
            f.write(c)
    return os.path.getsize(path)


def _put_side_info(p: _Paths, points_numbers, cube_positions):
    """The two files both modes share: point counts and cube positions.  -> (bytes_pointnums, bytes_cubepos)"""
    n_pn = _put(p.pointnums, np.asarray(points_numbers).astype(_U16).tobytes())
    pos = np.asarray(cube_positions)
    if pos.size and (pos.min() < 0 or pos.max() > 255):
        raise ValueError("cube positions must fit uint8 (the container stores them as bytes)")
    write_ply_data(p.cubepos_ply, pos.astype("uint8"))
    gpcc_encode(p.cubepos_ply, p.cubepos)
    return n_pn, os.path.getsize(p.cubepos)


def _get_side_info(p: _Paths):
    with open(p.pointnums, "rb") as f:
        points_numbers = np.frombuffer(f.read(), dtype=_U16)
    gpcc_decode(p.cubepos, p.cubepos_ply)
    return points_numbers, load_ply_data(p.cubepos_ply)


def _summary(**sizes):
    print("container: %d bytes  (%s)" % (sum(sizes.values()), ", ".join("%s %d" % kv for kv in sizes.items())))


# ------------------------------------------------------------------ factorized mode
def write_binary_files_factorized(filename, strings, points_numbers, cube_positions, min_v, max_v, shape, rootdir='./'):
    """-> (bytes_strings, bytes_pointnums, bytes_cubepos)"""
    os.makedirs(rootdir, exist_ok=True)
    p = _paths(filename, rootdir)
    n_str = _put(p.strings, _shape5(shape), _int8_pair(min_v, max_v, "latent"), _blob(strings))
    n_pn, n_pos = _put_side_info(p, points_numbers, cube_positions)
    _summary(strings=n_str, pointnums=n_pn, cubepos=n_pos)
    return n_str, n_pn, n_pos


def read_binary_files_factorized(filename, rootdir='./'):
    """-> (strings, points_numbers, cube_positions, min_v, max_v, shape)"""
    p = _paths(filename, rootdir)
    with open(p.strings, "rb") as f:
        raw = f.read()
    shape = np.frombuffer(raw, dtype=_I16, count=5)
    min_v, max_v = np.frombuffer(raw, dtype=_I8, count=2, offset=10)
    points_numbers, cube_positions = _get_side_info(p)
    return raw[12:], points_numbers, cube_positions, min_v, max_v, shape


# ------------------------------------------------------------------ hyperprior mode
def _pack_lengths(lens) -> bytes:
    out = bytearray()
    for n in lens:
        n = int(n)
        if 1 <= n <= 255:
            out.append(n)
        else:                                   # 0 and >= 256 take the escape form: a zero byte, then 16 bits
            out.append(0)
            out += np.array(n, dtype=_U16).tobytes()
    return bytes(out)


def _unpack_lengths(buf, pos, count):
    lens = np.empty(count, dtype=np.int32)
    for i in range(count):
        n = buf[pos]
        pos += 1
        if n == 0:
            n = int(np.frombuffer(buf, dtype=_U16, count=1, offset=pos)[0])
            pos += 2
        lens[i] = n
    return lens, pos


def write_binary_files_hyper(filename, y_strings, z_strings, points_numbers, cube_positions, y_min_vs, y_max_vs, y_shape,
                             z_min_v, z_max_v, z_shape, rootdir='./'):
    """-> (bytes_strings, bytes_strings_head, bytes_strings_hyper, bytes_pointnums, bytes_cubepos)"""
    os.makedirs(rootdir, exist_ok=True)
    p = _paths(filename, rootdir)
    ys = [bytes(s) for s in (y_strings.numpy() if hasattr(y_strings, "numpy") else y_strings)]
    lo = np.asarray(y_min_vs).astype(np.int64).reshape(-1)
    hi = np.asarray(y_max_vs).astype(np.int64).reshape(-1)
    lens = [len(s) for s in ys]
    if len(ys) > 32767:
        raise ValueError("%d cubes exceed the int16 cube count of the header" % len(ys))
    if ys and (lo.min() < -15 or lo.max() > 0 or hi.min() < 0 or hi.max() > 15):
        raise ValueError("y symbol ranges must lie in [-15, 0] / [0, 15]: they are packed into one byte as max*16 - min")
    if ys and max(lens) > 65535:
        raise ValueError("a y string of %d bytes exceeds the 16-bit length field" % max(lens))
    z_range = _int8_pair(z_min_v, z_max_v, "hyper-latent")
    n_head = _put(p.head, np.array(len(ys), dtype=_I16).tobytes(), (hi * 16 - lo).astype(_U8).tobytes(), _pack_lengths(lens),
                  _shape5(y_shape))
    n_str = _put(p.strings, b"".join(ys))
    n_hyp = _put(p.hyper, _shape5(z_shape), z_range, _blob(z_strings))
    n_pn, n_pos = _put_side_info(p, points_numbers, cube_positions)
    _summary(strings=n_str, strings_head=n_head, strings_hyper=n_hyp, pointnums=n_pn, cubepos=n_pos)
    return n_str, n_head, n_hyp, n_pn, n_pos


def read_binary_files_hyper(filename, rootdir='./'):
    """-> (y_strings, z_strings, points_numbers, cube_positions, y_min_vs, y_max_vs, y_shape, z_min_v, z_max_v, z_shape)"""
    p = _paths(filename, rootdir)
    with open(p.head, "rb") as f:
        head = f.read()
    count = int(np.frombuffer(head, dtype=_I16, count=1)[0])
    packed = np.frombuffer(head, dtype=_U8, count=count, offset=2).astype("int32")
    y_max_vs, y_min_vs = packed // 16, -(packed % 16)
    lens, pos = _unpack_lengths(head, 2 + count, count)
    y_shape = np.frombuffer(head, dtype=_I16, count=5, offset=pos)
    with open(p.strings, "rb") as f:
        body = f.read()
    ends = np.cumsum(lens)
    y_strings = np.empty(count, dtype=object)
    for i in range(count):
        y_strings[i] = body[ends[i] - lens[i]:ends[i]]
    with open(p.hyper, "rb") as f:
        raw = f.read()
    z_shape = np.frombuffer(raw, dtype=_I16, count=5)
    z_min_v, z_max_v = np.frombuffer(raw, dtype=_I8, count=2, offset=10)
    points_numbers, cube_positions = _get_side_info(p)
    return y_strings, raw[12:], points_numbers, cube_positions, y_min_vs, y_max_vs, y_shape, z_min_v, z_max_v, z_shape
