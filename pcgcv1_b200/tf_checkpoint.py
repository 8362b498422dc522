"""TF-1.13 checkpoint import: reads the TensorBundle files that
``tf.train.Checkpoint.save`` wrote for the reference (``train_hyper.py:106-121,268``; restored at ``transform.py:36-38,
107-112``) and maps the object-graph variable names onto this build's ``weights.npz`` keys, so the published pretrained
models run on the new path without TensorFlow.

Formats restated here (TensorFlow is not installable offline, so this reader is **parity unpinned**: no checkpoint of the
reference is available in this container; the round trip through ``write_bundle`` below is what the tests pin):

* ``<prefix>.index`` is a LevelDB-style sorted string table written WITHOUT compression: data blocks of prefix-compressed
  entries ``varint shared | varint non_shared | varint value_len | key suffix | value`` followed by a uint32 restart array
  and its length; every block is followed by a 5-byte trailer (type, masked crc32c); the 48-byte footer holds the
  metaindex and index block handles (varint offset, varint size) and the magic 0xdb4775248b80fb57.
* key ``""`` -> BundleHeaderProto; key = variable name -> BundleEntryProto {1: dtype, 2: shape{2: dim{1: size}},
  3: shard_id, 4: offset, 5: size, 6: crc32c}; tensor bytes live in ``<prefix>.data-SSSSS-of-NNNNN``, little endian.
* object-based checkpoints name a variable by its attribute path + ``/.ATTRIBUTES/VARIABLE_VALUE``, e.g.
  ``analysis_transform/vrn1_1/conv1_1/kernel/.ATTRIBUTES/VARIABLE_VALUE`` (attributes of model_voxception.py:21-54,
  83-122) and ``estimator/matrix_0/...`` (entropy_model.py:51-66).
"""
from __future__ import annotations

import glob
import os
import re
import struct
from typing import Dict, List, Tuple

import numpy as np

MAGIC = 0xDB4775248B80FB57
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_}
TOPS = ("analysis_transform", "synthesis_transform", "hyper_encoder", "hyper_decoder", "estimator")


# ---------------------------------------------------------------- crc32c (Castagnoli), masked as LevelDB does
def _crc_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_CRC = _crc_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------- varints / protobuf subset
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _fields(buf: bytes):
    """Yields (field number, wire type, value) of one protobuf message."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for num, wt, v in _fields(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            for n2, _, v2 in _fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _fields(v2):
                        if n3 == 1:
                            size = _signed(v3)
                    e["shape"].append(size)
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif num == 7:
            e["slices"] += 1
    return e


# ---------------------------------------------------------------- sorted string table
def _read_block(data: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    body = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        want = struct.unpack("<I", data[offset + size + 1:offset + size + 5])[0]
        if masked_crc(data[offset:offset + size + 1]) != want:
            raise ValueError("checkpoint index: block checksum mismatch at offset %d" % offset)
    if ctype != 0:
        raise ValueError("checkpoint index: compressed block (type %d); TensorBundle writes uncompressed tables" % ctype)
    return body


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path: str) -> List[Tuple[bytes, bytes]]:
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != MAGIC:
        raise ValueError("%s is not a TensorBundle index (bad magic)" % path)
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)           # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        out += _block_entries(_read_block(data, boff, bsize))
    return out


def read_bundle(prefix: str, verify_tensors: bool = False) -> Dict[str, np.ndarray]:
    """All tensors of the bundle ``<prefix>.index`` + ``<prefix>.data-*`` by variable name."""
    entries = read_table(prefix + ".index")
    num_shards = 1
    tensors: Dict[str, np.ndarray] = {}
    shards: Dict[int, bytes] = {}
    for key, val in entries:
        if key == b"":
            for num, _, v in _fields(val):
                if num == 1:
                    num_shards = v
                elif num == 2 and v != 0:
                    raise ValueError("big-endian TensorBundle is not supported")
            continue
        e = _parse_entry(val)
        if e["slices"]:
            raise ValueError("partitioned variable %r is not supported" % key.decode())
        if e["dtype"] not in DTYPES:
            continue                                  # strings (the object graph proto) and other non-numeric entries
        sid = e["shard_id"]
        if sid not in shards:
            with open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb") as f:
                shards[sid] = f.read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if verify_tensors and e["crc32c"] is not None and masked_crc(raw) != e["crc32c"]:
            raise ValueError("tensor %r: checksum mismatch" % key.decode())
        tensors[key.decode()] = np.frombuffer(raw, dtype=DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return tensors


def latest_checkpoint(ckpt_dir: str):
    """tf.train.latest_checkpoint: the prefix named by ``<ckpt_dir>/checkpoint``, else the highest-numbered ``*.index``."""
    state = os.path.join(ckpt_dir, "checkpoint")
    if os.path.exists(state):
        with open(state) as f:
            m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', f.read())
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(ckpt_dir, p)
            if os.path.exists(p + ".index"):
                return p
    idx = glob.glob(os.path.join(ckpt_dir, "*.index"))
    if not idx:
        return None
    num = lambda s: [int(x) for x in re.findall(r"\d+", os.path.basename(s))] or [0]
    return sorted(idx, key=num)[-1][:-len(".index")]


# ---------------------------------------------------------------- name mapping
def _layer_name(top: str, path: List[str]) -> str:
    """Attribute path inside a model -> the Keras ``name=`` this build keys weights by.  The two differ where the reference
    reuses attribute names: SynthesisTransform.vrnX_Y holds layers named dvrnX_Y_* (model_voxception.py:160-186) and
    HyperDecoder.convN holds layers named deconvN (:263-297)."""
    if len(path) == 2:                                                   # <vrn block attribute>/<conv attribute>
        block = path[0]
        if top == "synthesis_transform" and block.startswith("vrn"):
            block = "d" + block
        return block + "_" + path[1]
    layer = path[0]
    if top == "hyper_decoder" and re.fullmatch(r"conv\d(_\d)?", layer):
        layer = "de" + layer
    return layer


def map_names(tensors: Dict[str, np.ndarray], model: str = "voxception") -> Dict[str, np.ndarray]:
    """Object-graph variable names -> ``<net>/<keras layer name>/<kernel|bias>`` and ``estimator/<matrix_i|bais_i|factor_i>``.
    Optimizer slots, the global step and bookkeeping entries are dropped; an unrecognised model variable is an error."""
    out: Dict[str, np.ndarray] = {}
    unknown = []
    for name, arr in tensors.items():
        if not name.endswith(SUFFIX) or "/.OPTIMIZER_SLOT/" in name:
            continue
        parts = name[:-len(SUFFIX)].split("/")
        top = parts[0]
        if top not in TOPS:
            continue                                  # global_step, main_optimizer, save_counter
        if top == "estimator":
            m = re.fullmatch(r"(matrix|bais|bias|factor)_(\d)", parts[-1])
            if m:
                out["estimator/%s_%s" % ("bais" if m.group(1) in ("bais", "bias") else m.group(1), m.group(2))] = arr
                continue
            m = re.fullmatch(r"_(matrices|biases|factors)", parts[1]) if len(parts) == 3 else None
            if m and parts[2].isdigit():
                out["estimator/%s_%s" % ({"matrices": "matrix", "biases": "bais", "factors": "factor"}[m.group(1)], parts[2])] = arr
                continue
            unknown.append(name)
            continue
        var = parts[-1]
        if var not in ("kernel", "bias"):
            unknown.append(name)
            continue
        path = parts[1:-1]
        if len(path) not in (1, 2):
            unknown.append(name)
            continue
        out["%s/%s/%s" % (top, _layer_name(top, path), var)] = arr
    if unknown:
        raise KeyError("checkpoint variables this importer cannot place: %s" % ", ".join(sorted(unknown)[:8]))
    return out


def import_checkpoint(ckpt_dir: str, model: str = "voxception") -> Dict[str, np.ndarray]:
    prefix = latest_checkpoint(ckpt_dir)
    if prefix is None:
        raise FileNotFoundError("no TensorFlow checkpoint (*.index) in %s" % ckpt_dir)
    return map_names(read_bundle(prefix), model)


# ---------------------------------------------------------------- writer (tests, and exporting weights for the reference)
def _block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _msg(fields) -> bytes:
    out = bytearray()
    for num, wt, v in fields:
        out += _put_varint((num << 3) | wt)
        if wt == 0:
            out += _put_varint(v)
        elif wt == 2:
            out += _put_varint(len(v)) + v
        elif wt == 5:
            out += struct.pack("<I", v)
    return bytes(out)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_entries: int = 7) -> None:
    """Writes ``<prefix>.index`` / ``.data-00000-of-00001`` in the TensorBundle layout (keys are used as given)."""
    inv = {np.dtype(v): k for k, v in DTYPES.items()}
    data = bytearray()
    items: List[Tuple[bytes, bytes]] = [(b"", _msg([(1, 0, 1), (2, 0, 0), (3, 2, _msg([(1, 0, 1)]))]))]
    for name in sorted(tensors, key=lambda s: s.encode()):
        a = np.ascontiguousarray(tensors[name])
        raw = a.tobytes()
        shape = _msg([(2, 2, _msg([(1, 0, int(d))])) for d in a.shape])
        items.append((name.encode(), _msg([(1, 0, inv[a.dtype]), (2, 2, shape), (3, 0, 0), (4, 0, len(data)), (5, 0, len(raw)),
                                           (6, 5, masked_crc(raw))])))
        data += raw
    table = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(table)
        table.extend(block + b"\x00")
        table.extend(struct.pack("<I", masked_crc(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    index = []
    for i in range(0, len(items), block_entries):
        chunk = items[i:i + block_entries]
        index.append((chunk[-1][0], emit(_block(chunk))))
    meta = emit(_block([]))
    idx = emit(_block(index, 1))
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    table.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(table))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


def export_checkpoint(ckpt_dir: str, weights: Dict[str, np.ndarray], step: int = 1) -> str:
    """``weights.npz``-style dict -> an object-graph-named TensorBundle + ``checkpoint`` state file in ``ckpt_dir``."""
    os.makedirs(ckpt_dir, exist_ok=True)
    named = {}
    for k, v in weights.items():
        parts = k.split("/")
        if parts[0] == "estimator":
            named[k + SUFFIX] = v
            continue
        top, layer, var = parts
        m = re.fullmatch(r"d?(vrn\d_\d)_(conv\d_\d)", layer)
        path = "%s/%s" % (m.group(1), m.group(2)) if m else layer
        if top == "hyper_decoder" and layer.startswith("deconv"):
            path = layer[2:]
        named["%s/%s/%s%s" % (top, path, var, SUFFIX)] = v
    prefix = os.path.join(ckpt_dir, "ckpt-%d" % step)
    write_bundle(prefix, named)
    with open(os.path.join(ckpt_dir, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "ckpt-%d"\nall_model_checkpoint_paths: "ckpt-%d"\n' % (step, step))
    return prefix
