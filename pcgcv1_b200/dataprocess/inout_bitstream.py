"""Drop-in for ``dataprocess/inout_bitstream.py``: the on-disk container the CLI round-trips (SURVEY.md Appendix B,
section 8(f)).  Same five files with the same byte layout, so streams written by either side read on the other:

  <name>.strings        factorized: int16 shape[5] | int8 (min_v, max_v) | string      hyper: the y strings back to back
  <name>.strings_head   hyper only: int16 count | uint8 (max*16 - min) per cube | lengths (uint8, or 0 + int16 if > 255) | int16 y_shape[5]
  <name>.strings_hyper  hyper only: int16 z_shape[5] | int8 (z_min_v, z_max_v) | z string
  <name>.pointnums      uint16 per cube
  <name>.cubepos        cube positions through myutils.gpcc_wrapper (+ temp <name>_cubepos.ply)

The format's own limits are kept and checked instead of silently wrapping: <= 32767 cubes, y symbol range inside
[-15, 15] with max_v >= 0 >= min_v, string lengths < 65536 (the reference: < 32768), cube coordinates < 256."""
from __future__ import annotations

import os

import numpy as np

from ..myutils.gpcc_wrapper import gpcc_decode, gpcc_encode
from .inout_points import load_ply_data, write_ply_data


def _paths(filename, rootdir):
    j = lambda ext: os.path.join(rootdir, filename + ext)
    return j('.strings'), j('.strings_head'), j('.strings_hyper'), j('.pointnums'), j('.cubepos'), j('_cubepos.ply')


def _as_bytes(s):
    s = s.numpy() if hasattr(s, "numpy") else s
    if isinstance(s, np.ndarray):
        s = s.item() if s.dtype == object or s.ndim == 0 else s.tobytes()
    return bytes(s)


def _write_cubepos(cube_positions, ply_cubepos, file_cubepos):
    cube_positions = np.asarray(cube_positions)
    if cube_positions.size and (cube_positions.min() < 0 or cube_positions.max() > 255):
        raise ValueError("cube positions must fit uint8 (inout_bitstream.py:32,119 cast them)")
    write_ply_data(ply_cubepos, cube_positions.astype('uint8'))
    gpcc_encode(ply_cubepos, file_cubepos)


################### bitstream io without hyper prior ###################
def write_binary_files_factorized(filename, strings, points_numbers, cube_positions, min_v, max_v, shape, rootdir='./'):
    """inout_bitstream.py:10-44."""
    if not os.path.exists(rootdir):
        os.makedirs(rootdir)
    print('===== Write binary files =====')
    file_strings, _, _, file_pointnums, file_cubepos, ply_cubepos = _paths(filename, rootdir)
    min_v, max_v = int(np.asarray(min_v)), int(np.asarray(max_v))
    if not (-128 <= min_v <= 127 and -128 <= max_v <= 127):
        raise ValueError("symbol range [%d, %d] does not fit the int8 header" % (min_v, max_v))
    with open(file_strings, 'wb') as f:
        f.write(np.array(shape, dtype=np.int16).tobytes())      # [batch size, length, width, height, channels]
        f.write(np.array((min_v, max_v), dtype=np.int8).tobytes())
        f.write(_as_bytes(strings))
    with open(file_pointnums, 'wb') as f:
        f.write(np.array(points_numbers, dtype=np.uint16).tobytes())
    _write_cubepos(cube_positions, ply_cubepos, file_cubepos)
    bytes_strings = os.path.getsize(file_strings)
    bytes_pointnums = os.path.getsize(file_pointnums)
    bytes_cubepos = os.path.getsize(file_cubepos)
    print('Total file size (Bytes): {}'.format(bytes_strings + bytes_pointnums + bytes_cubepos))
    print('Strings (Bytes): {}'.format(bytes_strings))
    print('Numbers of points (Bytes): {}'.format(bytes_pointnums))
    print('Positions of cubes (Bytes): {}'.format(bytes_cubepos))
    return bytes_strings, bytes_pointnums, bytes_cubepos


def read_binary_files_factorized(filename, rootdir='./'):
    """inout_bitstream.py:46-70."""
    print('===== Read binary files =====')
    file_strings, _, _, file_pointnums, file_cubepos, ply_cubepos = _paths(filename, rootdir)
    with open(file_strings, 'rb') as f:
        shape = np.frombuffer(f.read(2 * 5), dtype=np.int16)
        min_v, max_v = np.frombuffer(f.read(1 * 2), dtype=np.int8)
        strings = f.read()
    with open(file_pointnums, 'rb') as f:
        points_numbers = np.frombuffer(f.read(), dtype=np.uint16)
    gpcc_decode(file_cubepos, ply_cubepos)
    cube_positions = load_ply_data(ply_cubepos)
    return strings, points_numbers, cube_positions, min_v, max_v, shape


################### bitstream io with hyper prior ###################
def write_binary_files_hyper(filename, y_strings, z_strings, points_numbers, cube_positions, y_min_vs, y_max_vs, y_shape,
                             z_min_v, z_max_v, z_shape, rootdir='./'):
    """inout_bitstream.py:75-141."""
    if not os.path.exists(rootdir):
        os.makedirs(rootdir)
    print('===== Write binary files =====')
    file_strings, file_strings_head, file_strings_hyper, file_pointnums, file_cubepos, ply_cubepos = _paths(filename, rootdir)
    y_strings = [bytes(s) for s in list(y_strings.numpy() if hasattr(y_strings, "numpy") else y_strings)]
    y_min_vs = np.asarray(y_min_vs).astype(np.int64).reshape(-1)
    y_max_vs = np.asarray(y_max_vs).astype(np.int64).reshape(-1)
    if len(y_strings) > 32767:
        raise ValueError("%d cubes exceed the int16 cube count of the header" % len(y_strings))
    if len(y_strings) and (y_min_vs.min() < -15 or y_min_vs.max() > 0 or y_max_vs.min() < 0 or y_max_vs.max() > 15):
        raise ValueError("y symbol ranges must lie in [-15, 0] / [0, 15] to pack as max*16 - min (inout_bitstream.py:95-96)")
    lens = np.array([len(s) for s in y_strings], dtype=np.int64)
    if len(lens) and lens.max() > 65535:
        raise ValueError("a y string of %d bytes exceeds the 16-bit length field" % int(lens.max()))
    z_min_v, z_max_v = int(np.asarray(z_min_v)), int(np.asarray(z_max_v))
    if not (-128 <= z_min_v <= 127 and -128 <= z_max_v <= 127):
        raise ValueError("z symbol range [%d, %d] does not fit the int8 header" % (z_min_v, z_max_v))
    with open(file_strings_head, 'wb') as f:
        f.write(np.array(len(y_strings), dtype=np.int16).tobytes())
        f.write(np.array(y_max_vs * 16 - y_min_vs, dtype=np.uint8).tobytes())
        head = bytearray()
        for l in lens:
            if 0 < l <= 255:
                head += np.array(l, dtype=np.uint8).tobytes()
            else:
                # the reference writes a bare 0 byte for an EMPTY string and then cannot read it back (a 0 byte announces
                # an int16); writing the escape form for l == 0 keeps the stream decodable and is what its reader expects
                # lengths 32768..65535 (random-init weights code ~40 KB per cube) use the same 16 bits unsigned; the reference
                # would overflow its int16 there, so no stream it can write is read differently
                head += np.array(0, dtype=np.uint8).tobytes() + np.array(l, dtype=np.uint16).tobytes()
        f.write(bytes(head))
        f.write(np.array(y_shape, dtype=np.int16).tobytes())   # [batch size, length, width, height, channels]
    with open(file_strings, 'wb') as f:
        f.write(b"".join(y_strings))
    with open(file_strings_hyper, 'wb') as f:
        f.write(np.array(z_shape, dtype=np.int16).tobytes())
        f.write(np.array((z_min_v, z_max_v), dtype=np.int8).tobytes())
        f.write(_as_bytes(z_strings))
    with open(file_pointnums, 'wb') as f:
        f.write(np.array(points_numbers, dtype=np.uint16).tobytes())
    _write_cubepos(cube_positions, ply_cubepos, file_cubepos)
    bytes_strings = os.path.getsize(file_strings)
    bytes_strings_head = os.path.getsize(file_strings_head)
    bytes_strings_hyper = os.path.getsize(file_strings_hyper)
    bytes_pointnums = os.path.getsize(file_pointnums)
    bytes_cubepos = os.path.getsize(file_cubepos)
    print('Total file size (Bytes): {}'.format(bytes_strings + bytes_strings_head + bytes_strings_hyper + bytes_pointnums + bytes_cubepos))
    print('Strings (Bytes): {}'.format(bytes_strings))
    print('Strings head (Bytes): {}'.format(bytes_strings_head))
    print('Strings hyper (Bytes): {}'.format(bytes_strings_hyper))
    print('Numbers of points (Bytes): {}'.format(bytes_pointnums))
    print('Positions of cubes (Bytes): {}'.format(bytes_cubepos))
    return bytes_strings, bytes_strings_head, bytes_strings_hyper, bytes_pointnums, bytes_cubepos


def read_binary_files_hyper(filename, rootdir='./'):
    """inout_bitstream.py:144-198."""
    print('===== Read binary files =====')
    file_strings, file_strings_head, file_strings_hyper, file_pointnums, file_cubepos, ply_cubepos = _paths(filename, rootdir)
    with open(file_strings_head, 'rb') as f:
        head = f.read()
    y_strings_num = int(np.frombuffer(head[:2], dtype=np.int16)[0])
    y_max_min_vs = np.frombuffer(head[2:2 + y_strings_num], dtype=np.uint8).astype('int32')
    y_max_vs = y_max_min_vs // 16
    y_min_vs = -(y_max_min_vs % 16)
    pos = 2 + y_strings_num
    y_strings_lens = np.empty(y_strings_num, dtype=np.int32)
    for i in range(y_strings_num):
        l = head[pos]
        pos += 1
        if l == 0:
            l = int(np.frombuffer(head[pos:pos + 2], dtype=np.uint16)[0])
            pos += 2
        y_strings_lens[i] = l
    y_shape = np.frombuffer(head[pos:pos + 10], dtype=np.int16)
    with open(file_strings, 'rb') as f:
        body = f.read()
    ends = np.cumsum(y_strings_lens)
    y_strings = np.empty(y_strings_num, dtype=object)
    for i in range(y_strings_num):
        y_strings[i] = body[ends[i] - y_strings_lens[i]:ends[i]]
    with open(file_strings_hyper, 'rb') as f:
        z_shape = np.frombuffer(f.read(2 * 5), dtype=np.int16)
        z_min_v, z_max_v = np.frombuffer(f.read(1 * 2), dtype=np.int8)
        z_strings = f.read()
    with open(file_pointnums, 'rb') as f:
        points_numbers = np.frombuffer(f.read(), dtype=np.uint16)
    gpcc_decode(file_cubepos, ply_cubepos)
    cube_positions = load_ply_data(ply_cubepos)
    return y_strings, z_strings, points_numbers, cube_positions, y_min_vs, y_max_vs, y_shape, z_min_v, z_max_v, z_shape
