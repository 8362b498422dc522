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


def decode_slices(n: int, world: int, z_ratio: Optional[float] = None) -> List[Tuple[int, int]]:
    """Contiguous slices for the DECODE side of the packed GPU path.  The hyper latents of all cubes are ONE sequential string
    (entropy_model.py:249-259) that rank 0 decodes at ``z_ratio`` times the rate at which one GPU decodes cubes, and rank r > 0
    can only start once the decoder has passed the END of its slice -- with count-balanced slices the last rank starts when the
    whole string is decoded and the cloud takes (string time + 1/world of the GPU time).  Slices that make every rank FINISH at
    the same time T instead: rank 0 (fed progressively from the head of the string) takes x = T * gpu_rate cubes, and
    S_r = (S_{r-1} + x) * z_ratio / (1 + z_ratio) are the slice ends of the ranks after it (arrival S_r / z_rate plus
    (S_r - S_{r-1}) / gpu_rate of work = T).  z_ratio -> infinity gives the balanced slices back.  The stream and the decoded
    cloud do not depend on the slicing: cubes are independent and results are gathered in rank order.
    ``z_ratio`` defaults to PCGC_SHARD_Z_RATIO (3.0: 0.35 s of string decoding against 0.99 s of GPU decoding for the 7 769-cube
    cloud on one B200; measured N = 2: 4 925 cubes/s balanced, 5 431 at 2.2, 5 467 at 3.0; N = 4: 7 476 balanced, 8 315 at 3.0,
    8 155 at 4.0); <= 0 selects the balanced slices."""
    if z_ratio is None:
        import os
        z_ratio = float(os.environ.get("PCGC_SHARD_Z_RATIO", "3.0"))
    if world <= 1 or n <= 0 or not (z_ratio > 0):
        return shard_slices(n, world)
    k = z_ratio / (1.0 + z_ratio)
    coef = [1.0]
    for _ in range(1, world):
        coef.append(k * (coef[-1] + 1.0))
    x = n / coef[-1]
    ends = [min(n, max(0, int(round(c * x)))) for c in coef]
    ends[-1] = n
    for r in range(1, world):
        ends[r] = max(ends[r], ends[r - 1])
    return [(0 if r == 0 else ends[r - 1], ends[r]) for r in range(world)]


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

    def decode_local_points(self, y_strings: Sequence[bytes], y_min, y_max, z_hat: np.ndarray, nums: np.ndarray, rho: float):
        """-> (points int16 [n,3] in np.where order per cube, counts int32 [b]); default: from decode_local's masks."""
        masks = np.asarray(self.decode_local(y_strings, y_min, y_max, z_hat, nums, rho))
        pts = [np.argwhere((m[..., 0] if m.shape[-1] == 1 else m) > 0) for m in masks]
        counts = np.array([len(p) for p in pts], np.int32)
        width = masks.ndim - 1 - (1 if masks.shape[-1] == 1 else 0)
        return (np.concatenate(pts).astype(np.int16) if pts else np.zeros((0, width), np.int16)), counts


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
        """The rank's slice through the SAME pipeline as transform.compress_hyper (chunked transforms, GPU range coder on the
        coder stream), minus the hyper string: z is one global string (entropy_model.py:249-259) and is coded on rank 0."""
        from . import transform
        c = self.codec
        if len(cubes) == 0:
            return {"y_strings": [], "y_min": np.zeros(0, np.int32), "y_max": np.zeros(0, np.int32), "z_hat": np.zeros((0, 8, 8, 8, 8), np.int16)}
        if self.runtime.coder_mode() == "gpu":
            strings, mm, z_all, _, _, _, _ = transform._compress_hyper_gpu_coder(c, self.eb, self.sc, cubes, False, code_z=False)
            return {"y_strings": strings, "y_min": mm[:, 0].astype(np.int32), "y_max": mm[:, 1].astype(np.int32),
                    "z_hat": self.runtime.to_host(z_all).astype(np.int16)}
        ys = c.analysis(c.to_device(cubes))
        z_hat, _, _, _ = c.factorized(self.eb._slot, c.hyper_encode(ys), want_p=False, want_bits=False)
        locs, scales = c.hyper_decode(z_hat, 1e-9)
        strings, mn, mx = self.sc.compress_cubes(ys, locs, scales)
        return {"y_strings": strings, "y_min": mn.astype(np.int32), "y_max": mx.astype(np.int32),
                "z_hat": self.runtime.to_host(z_hat).astype(np.int16)}

    def encode_z(self, z_hat_all):
        """The quantised hyper latents of ALL cubes -> one string with one (min_v, max_v) (entropy_model.py:246-259).  The values
        are integers already, so round() is the identity and floor(min) / ceil(max) are min / max: no GPU round trip of the
        (cubes x 4096) tensor, just the per-channel CDF table and the host range coder."""
        z = np.ascontiguousarray(z_hat_all)
        flat = z.reshape(-1)
        mn, mx = int(flat.min()), int(flat.max())
        c, slot = self.eb._resolve(z.shape[-1])
        cdf = c.factorized_cdf(slot, mn, mx, self.eb._likelihood_bound, self.eb._range_coder_precision)
        sym = np.ascontiguousarray(flat.astype(np.int16) - np.int16(mn))
        return self.runtime.range_encode(sym, cdf, self.eb._range_coder_precision), mn, mx

    def encode_z_async(self, z_hat_all):
        """encode_z with the host range coder (one sequential string, ~8 ns per symbol, GIL released) on a worker thread:
        -> join() returning (string, min_v, max_v).  The CDF table is built here, on the caller's thread."""
        z = np.ascontiguousarray(z_hat_all)
        flat = z.reshape(-1)
        mn, mx = int(flat.min()), int(flat.max())
        c, slot = self.eb._resolve(z.shape[-1])
        cdf = c.factorized_cdf(slot, mn, mx, self.eb._likelihood_bound, self.eb._range_coder_precision)
        sym = np.ascontiguousarray(flat.astype(np.int16) - np.int16(mn))
        from concurrent.futures import ThreadPoolExecutor
        ex = ThreadPoolExecutor(max_workers=1)
        job = ex.submit(self.runtime.range_encode, sym, cdf, self.eb._range_coder_precision)

        def join():
            out = job.result()
            ex.shutdown(wait=False)
            return out, mn, mx
        return join

    def decode_z(self, z_string, z_min, z_max, z_shape):
        dec, mn, _ = self.decode_z_begin(z_string, z_min, z_max, z_shape)
        return dec.finish().reshape([int(v) for v in z_shape]) + np.int16(mn)

    def decode_z_begin(self, z_string, z_min, z_max, z_shape):
        """-> (runtime.ProgressiveDecode, min_v, symbols per cube): the string decodes on a worker thread; ``dec.wait(k)`` returns
        once the first k symbols (value - min_v, int16) exist.  The string is sequential, so the first cubes come first."""
        shape = [int(v) for v in np.asarray(z_shape).reshape(-1)]
        mn, mx = int(np.asarray(z_min)), int(np.asarray(z_max))
        c, slot = self.eb._resolve(shape[-1])
        cdf = c.factorized_cdf(slot, mn, mx, self.eb._likelihood_bound, self.eb._range_coder_precision)
        per = int(np.prod(shape[1:]))
        return self.runtime.ProgressiveDecode(bytes(z_string), shape[0] * per, cdf, self.eb._range_coder_precision), mn, per

    # ---- device-resident strings (the sharded fast path: the per-cube strings of a slice never become Python objects on the
    # ranks; they move as ONE packed uint8 tensor, GPU -> GPU) -----------------------------------------------------------------
    packed_exchange = True

    def encode_local_packed(self, cubes):
        """-> dict as encode_local with 'y_packed' (uint8 device tensor) + 'y_lens' (int64 [b]) instead of 'y_strings'."""
        import torch
        from . import transform
        c = self.codec
        if len(cubes) == 0:
            return {"y_packed": torch.zeros(0, dtype=torch.uint8, device=c.dev), "y_lens": np.zeros(0, np.int64), "y_min": np.zeros(0, np.int32),
                    "y_max": np.zeros(0, np.int32), "z_hat": np.zeros((0, 8, 8, 8, 8), np.int16)}
        (packed, off), mm, z_all, _, _, _, _ = transform._compress_hyper_gpu_coder(c, self.eb, self.sc, cubes, False, code_z=False,
                                                                                  strings_on_device=True)
        return {"y_packed": packed, "y_lens": np.diff(off), "y_min": mm[:, 0].astype(np.int32), "y_max": mm[:, 1].astype(np.int32),
                "z_hat": self.runtime.to_host(z_all).astype(np.int16)}

    def _decode_masks_dev(self, y_strings, y_min, y_max, z_hat, nums, rho, uploaded=None):
        """-> uint8 mask tensor [b,64,64,64,1] on the device (transform.decompress_hyper's pipeline + top-k).  ``uploaded`` =
        (packed uint8 device tensor, offsets int64 device tensor [b+1]) when the strings are already in HBM."""
        import torch
        from . import transform
        c = self.codec
        b = len(y_min)
        if callable(z_hat):
            z_get = z_hat                                                    # (a, e) -> float32 device tensor [e-a,8,8,8,8]; may block
        else:
            zd = c.to_device(np.ascontiguousarray(z_hat, dtype=np.float32))
            z_get = lambda a, e: zd[a:e]
        if self.runtime.coder_mode() == "gpu":
            xs = transform._decompress_hyper_gpu_coder(c, self.sc, None if uploaded is not None else list(y_strings), np.asarray(y_min),
                                                       np.asarray(y_max), [1, 16, 16, 16, 16], z_get,
                                                       transform._gpu_decode_chunks(b), uploaded=uploaded)
        else:
            locs, scales = c.hyper_decode(z_get(0, b), 1e-9)
            ys = self.sc.decompress_cubes(list(y_strings), locs, scales, y_min, y_max)
            xs = c.synthesis(ys.reshape(b, 16, 16, 16, 16))
        ks = np.array([int(rho * np.array(n)) for n in nums], np.int32)
        mask, _, _ = c.topk(xs, c.to_device(ks))
        return mask

    def decode_local(self, y_strings, y_min, y_max, z_hat, nums, rho):
        if len(y_strings) == 0:
            return np.zeros((0, 64, 64, 64, 1), np.uint8)
        return self.runtime.to_host(self._decode_masks_dev(y_strings, y_min, y_max, z_hat, nums, rho), "mask").copy()

    def decode_local_points(self, y_strings, y_min, y_max, z_hat, nums, rho, uploaded=None):
        """-> (points int16 [n,3], counts int32 [b]): the decoder's final product (voxels2points, inout_points.py:134-143) taken on
        the device, 6 bytes per point instead of 256 KiB of mask per cube across PCIe and the host gather."""
        if len(y_min) == 0:
            return np.zeros((0, 3), np.int16), np.zeros(0, np.int32)
        mask = self._decode_masks_dev(y_strings, y_min, y_max, z_hat, nums, rho, uploaded=uploaded)
        cap = int(np.asarray(nums, np.int64).sum() * max(1.0, rho) * 1.25) + 1024
        return self.codec.extract_points(mask, cap)


def _dist():
    import torch.distributed as dist
    return dist


def _packed_ok(local, group) -> bool:
    """The packed GPU -> GPU exchange needs a CUDA local codec and, for more than one rank, an NCCL default group."""
    dist = _dist()
    if not getattr(local, "packed_exchange", False) or local.runtime.coder_mode() != "gpu":
        return False
    return dist.get_world_size(group) == 1 or dist.get_backend() == "nccl"


def compress_sharded(cubes_local: np.ndarray, local: LocalCodec, group=None) -> Optional[dict]:
    """Every rank passes ITS slice (``shard_slices`` order).  Rank 0 returns the stream of the whole cloud
    {'y_strings', 'y_min', 'y_max', 'y_shape', 'z_string', 'z_min', 'z_max', 'z_shape'}; other ranks None.

    With a GPU codec the per-cube strings of a slice stay packed in HBM and reach rank 0 as ONE uint8 tensor per rank (NCCL
    send / recv on the default group; ``group`` carries the small host objects): the stream then also holds 'y_blob' (all
    strings back to back, NumPy uint8) + 'y_lens', and 'y_strings' is a lazily sliced view of them."""
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if _packed_ok(local, group):
        import torch
        part = local.encode_local_packed(cubes_local)
        packed = part.pop("y_packed")
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0, group=group)
        if rank != 0:
            if packed.numel():
                dist.send(packed, dst=0)
            return None
        z_hat = np.concatenate([p["z_hat"] for p in parts])
        z_join = local.encode_z_async(z_hat)                                  # the one hyper string, coded beside the string gather
        totals = [int(p["y_lens"].sum()) for p in parts]
        dev_parts = [packed]
        for r in range(1, world):
            buf = torch.empty(totals[r], dtype=torch.uint8, device=packed.device)
            if totals[r]:
                dist.recv(buf, src=r)
            dev_parts.append(buf)
        allp = torch.cat(dev_parts) if world > 1 else packed
        blob = local.runtime.to_host(allp, "stream_blob").copy() if allp.numel() else np.zeros(0, np.uint8)
        lens = np.concatenate([p["y_lens"] for p in parts]).astype(np.int64)
        z_string, z_min, z_max = z_join()
        return {"y_strings": BlobStrings(blob, lens), "y_blob": blob, "y_lens": lens, "y_min": np.concatenate([p["y_min"] for p in parts]),
                "y_max": np.concatenate([p["y_max"] for p in parts]), "y_shape": np.array([1, 16, 16, 16, 16], np.int64),
                "z_string": z_string, "z_min": z_min, "z_max": z_max, "z_shape": np.array(z_hat.shape, np.int32)}
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


class BlobStrings:
    """Read-only list of byte strings backed by one uint8 blob + lengths (7769 strings of 40 KB are sliced on demand)."""

    def __init__(self, blob: np.ndarray, lens: np.ndarray):
        self.blob, self.lens = blob, np.asarray(lens, np.int64)
        self.off = np.zeros(len(self.lens) + 1, np.int64)
        np.cumsum(self.lens, out=self.off[1:])

    def __len__(self):
        return len(self.lens)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self.blob[self.off[i]:self.off[i + 1]].tobytes()

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __eq__(self, other):
        return list(self) == list(other)


def decompress_sharded(stream: Optional[dict], nums: Optional[np.ndarray], rho: float, local: LocalCodec, group=None, output: str = "masks"):
    """Rank 0 passes the stream (+ the per-cube point counts); every rank decodes its slice; rank 0 returns the
    uint8 occupancy masks of all cubes in order (``output="masks"``) or ``(points int16 [n,3], counts int32 [B])``
    (``output="points"``: what the decoder writes to the .ply; a few MB through the gather instead of 256 KiB per cube)."""
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if _packed_ok(local, group):
        import torch
        c = local.codec
        pieces, up, off_all, slices, dec = None, None, None, None, None
        if rank == 0:
            B = len(stream["y_min"])
            # the ONE hyper string starts decoding on a worker thread now; everything below overlaps it
            dec, z_min, per = local.decode_z_begin(stream["z_string"], stream["z_min"], stream["z_max"], stream["z_shape"])
            if "y_blob" in stream:
                blob, lens = stream["y_blob"], np.asarray(stream["y_lens"], np.int64)
            else:
                ys = list(stream["y_strings"])
                lens = np.fromiter((len(x) for x in ys), np.int64, len(ys))
                blob = np.frombuffer(b"".join(ys), np.uint8)
            off_all = np.zeros(B + 1, np.int64)
            np.cumsum(lens, out=off_all[1:])
            slices = decode_slices(B, world)                                  # skewed: every rank finishes together (see decode_slices)
            pieces = [{"y_lens": lens[a:b], "y_min": stream["y_min"][a:b], "y_max": stream["y_max"][a:b], "z_min": z_min, "z_per": per,
                       "z_tail": [int(v) for v in np.asarray(stream["z_shape"]).reshape(-1)[1:]], "nums": np.asarray(nums)[a:b]} for a, b in slices]
            up = c.to_device(blob) if len(blob) else torch.zeros(0, dtype=torch.uint8, device=c.dev)      # ONE H2D copy of the whole stream
        mine = [None]
        dist.scatter_object_list(mine, pieces, src=0, group=group)
        m = mine[0]
        total = int(m["y_lens"].sum())
        if rank == 0:
            torch.cuda.current_stream(c.dev).synchronize()
            for r in range(1, world):
                a, b = slices[r]
                if off_all[b] > off_all[a]:
                    dist.send(up[off_all[a]:off_all[b]].contiguous(), dst=r)
            packed = up[:off_all[slices[0][1]]]
        else:
            packed = torch.empty(total, dtype=torch.uint8, device=c.dev)
            if total:
                dist.recv(packed, src=0)
        off = np.zeros(len(m["y_lens"]) + 1, np.int64)
        np.cumsum(m["y_lens"], out=off[1:])
        uploaded = (packed if packed.numel() else torch.zeros(1, dtype=torch.uint8, device=c.dev), c.to_device(off))
        per, z_min, z_tail = m["z_per"], m["z_min"], m["z_tail"]
        nb = len(m["y_min"])
        sender = None
        host_z = world == 1 or dist.get_backend(group) == "gloo"         # int16 symbols travel over the host group when there is one
        if rank == 0:
            a0 = slices[0][0]

            def z_get(a, e):                                                 # blocks until cubes [a, e) of rank 0's slice are decoded
                sym = dec.wait((a0 + e) * per)[(a0 + a) * per:(a0 + e) * per]
                return c.to_device((sym.astype(np.float32) + np.float32(z_min)).reshape([e - a] + z_tail))

            if world > 1:
                # the other ranks' hyper latents leave as soon as the sequential decoder reaches the end of their slice (int16
                # symbols, 8 KB per cube, host group), on a helper thread: this thread is busy feeding rank 0's own GPU
                def ship():
                    for r in range(1, world):
                        a, b = slices[r]
                        if b > a:
                            t = torch.from_numpy(dec.wait(b * per)[a * per:b * per].copy())
                            if host_z:
                                dist.send(t, dst=r, group=group)
                            else:
                                dist.send(t.to(c.dev), dst=r)
                import threading
                sender = threading.Thread(target=ship, daemon=True)
                sender.start()
        else:
            zs = torch.empty(nb * per, dtype=torch.int16, device=None if host_z else c.dev)
            if nb:
                if host_z:
                    dist.recv(zs, src=0, group=group)
                else:
                    dist.recv(zs, src=0)
            zd = c.to_device((zs.cpu().numpy().astype(np.float32) + np.float32(z_min)).reshape([nb] + z_tail)) if nb else None
            z_get = lambda a, e: zd[a:e]
        if output == "points":
            res = local.decode_local_points(None, m["y_min"], m["y_max"], z_get, m["nums"], rho, uploaded=uploaded)
        else:
            res = (local.runtime.to_host(local._decode_masks_dev(None, m["y_min"], m["y_max"], z_get, m["nums"], rho, uploaded=uploaded), "mask").copy()
                   if nb else np.zeros((0, 64, 64, 64, 1), np.uint8))
        if sender is not None:
            sender.join()
        parts = [None] * world if rank == 0 else None
        dist.gather_object(res, parts, dst=0, group=group)
        if rank != 0:
            return None
        if output == "points":
            return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        return np.concatenate(parts)
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
    if output == "points":
        res = local.decode_local_points(m["y_strings"], m["y_min"], m["y_max"], m["z_hat"], m["nums"], rho)
    else:
        res = local.decode_local(m["y_strings"], m["y_min"], m["y_max"], m["z_hat"], m["nums"], rho)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(res, parts, dst=0, group=group)
    if rank != 0:
        return None
    if output == "points":
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    return np.concatenate(parts)
