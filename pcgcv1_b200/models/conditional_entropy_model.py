"""Drop-in for ``models/conditional_entropy_model.py``: ``SymmetricConditional`` -- the per-element
LAPLACE model conditioned on the hyper decoder's (loc, scale) (conditional_entropy_model.py:21-201;
north_star calls it "Gaussian", the code is Laplace and that is what runs).

``compress``/``decompress`` keep the reference semantics (the whole tensor is ONE string with one
symbol range); ``compress_cubes``/``decompress_cubes`` are the batched form of the per-cube loops in
``transform.py:157-168,238-247``: one string and one (min_v, max_v) per cube, CDF rows built on the
GPU for all cubes in one launch and the strings coded on a host thread pool."""
from __future__ import annotations

import numpy as np
import torch

from .. import runtime


# Per-cube symbol ranges are widened to contain 0 and at least two symbols before coding (pcgc_widen_symbol_ranges): a cube whose
# latents are all zero, or all on one side of zero, would otherwise abort the whole cloud after all GPU work (one-symbol alphabet
# / a range the container's header byte cannot hold; the reference has the same limitation).  The ranges the reference would
# produce are unchanged whenever they already straddle zero with two symbols, i.e. for every cube seen in the test workloads.
# PCGC_WIDEN_RANGES=0 restores the reference's exact (min, max).
WIDEN_RANGES = bool(int(__import__("os").environ.get("PCGC_WIDEN_RANGES", "1")))


class SymmetricConditional:
    def __init__(self, likelihood_bound=1e-9, range_coder_precision=16, codec=None):
        self._likelihood_bound = float(likelihood_bound)
        self._range_coder_precision = int(range_coder_precision)
        if self._range_coder_precision != 16:
            raise NotImplementedError("the CDF kernels emit 16-bit tables (the reference's only setting)")
        self._codec = codec

    def bind(self, codec):
        self._codec = codec
        return self

    @property
    def codec(self):
        if self._codec is None:
            self._codec = runtime.get_codec("voxception", "")
        return self._codec

    def _dev3(self, inputs, loc, scale):
        c = self.codec
        return (c.to_device(inputs, torch.float32), c.to_device(loc, torch.float32), c.to_device(scale, torch.float32))

    def __call__(self, inputs, loc, scale, training=False, seed=0):
        """-> (quantised inputs, max(likelihood, bound)) (conditional_entropy_model.py:71-93).  ``training=True`` is the
        reference's "noise" mode (:62-64): inputs + U(-1/2, 1/2) from a Philox stream keyed by ``seed`` (the reference is
        unseeded), likelihood evaluated at the noisy value."""
        y, l, s = self._dev3(inputs, loc, scale)
        c = self.codec
        c.set_quantize_mode(bool(training), seed)
        try:
            y_hat, p, _, _ = c.laplace(y.reshape(1, -1), l.reshape(1, -1), s.reshape(1, -1), self._likelihood_bound,
                                       want_p=True, want_bits=False)
        finally:
            c.set_quantize_mode(False)
        return runtime.DeviceResult(y_hat.reshape(y.shape)), runtime.DeviceResult(p.reshape(y.shape))

    def estimate_bits(self, inputs, loc, scale) -> np.ndarray:
        """per leading-index bits: sum(log2 p) * -1 (train_hyper.py:148-150)."""
        y, l, s = self._dev3(inputs, loc, scale)
        B = y.shape[0]
        _, _, bits, _ = self.codec.laplace(y.reshape(B, -1), l.reshape(B, -1), s.reshape(B, -1), self._likelihood_bound,
                                           want_p=False, want_bits=True)
        return runtime.to_host(bits)

    # ---- reference semantics: one string for the whole tensor -----------------------------------
    def compress(self, inputs, loc, scale):
        y, l, s = self._dev3(inputs, loc, scale)
        strings, mins, maxs = self.compress_cubes(y.reshape(1, -1), l.reshape(1, -1), s.reshape(1, -1))
        return runtime.HostResult(strings[0]), runtime.HostResult(np.int32(mins[0])), runtime.HostResult(np.int32(maxs[0]))

    def decompress(self, strings, loc, scale, min_v, max_v, datashape):
        strings = runtime.unwrap(strings)
        if isinstance(strings, np.ndarray):
            strings = strings.item() if strings.ndim == 0 else strings[0]
        datashape = [int(v) for v in np.asarray(runtime.unwrap(datashape)).reshape(-1)]
        c = self.codec
        l = c.to_device(loc, torch.float32).reshape(1, -1)
        s = c.to_device(scale, torch.float32).reshape(1, -1)
        y = self.decompress_cubes([strings], l, s, [int(np.asarray(runtime.unwrap(min_v)))], [int(np.asarray(runtime.unwrap(max_v)))])
        return runtime.DeviceResult(y.reshape(datashape))

    # ---- batched per-cube form -------------------------------------------------------------------
    def compress_cubes(self, ys, locs, scales, threads: int = 0):
        """ys/locs/scales torch float32 [B, ...] on the device -> (list of B strings, min_vs[B], max_vs[B])."""
        c = self.codec
        B = ys.shape[0]
        y2, l2, s2 = ys.reshape(B, -1), locs.reshape(B, -1), scales.reshape(B, -1)
        y_hat, _, _, mm = c.laplace(y2, l2, s2, self._likelihood_bound, want_p=False, want_bits=False)
        if WIDEN_RANGES:
            c.widen_symbol_ranges(mm)
        iv = c.laplace_intervals(y_hat, l2, s2, mm, self._likelihood_bound)
        mm_h = runtime.to_host(mm)
        if runtime.coder_mode() == "gpu" and B > 0:
            packed, offsets = self.encode_dev(iv)
            off = runtime.to_host(offsets)
            blob = runtime.to_host(packed[:int(off[B])], "enc_bytes")
            c.synchronize()
            strings = [blob[off[i]:off[i + 1]].tobytes() for i in range(B)]
        else:
            strings = runtime.range_encode_intervals_batch(runtime.to_host(iv, "intervals"), threads)
        return strings, mm_h[:, 0].copy(), mm_h[:, 1].copy()

    # ---- split forms for the software pipeline in transform.py: GPU part now, host coder part later ---------------
    def encode_begin(self, ys, locs, scales, slot: int):
        """GPU half of compress_cubes: returns (pinned intervals, copy-done event, minmax host array)."""
        c = self.codec
        B = ys.shape[0]
        y2, l2, s2 = ys.reshape(B, -1), locs.reshape(B, -1), scales.reshape(B, -1)
        y_hat, _, _, mm = c.laplace(y2, l2, s2, self._likelihood_bound, want_p=False, want_bits=False)
        if WIDEN_RANGES:
            c.widen_symbol_ranges(mm)
        iv = c.laplace_intervals(y_hat, l2, s2, mm, self._likelihood_bound)
        stage, done = runtime.to_host_async(iv, "intervals%d" % slot)
        return stage, done, runtime.to_host(mm)

    @staticmethod
    def encode_finish(stage, done, threads: int = 0):
        done.synchronize()
        return runtime.range_encode_intervals_batch(stage.numpy(), threads)

    def decode_begin(self, locs, scales, min_vs, max_vs, slot: int):
        """GPU half of decompress_cubes: per-element CDF rows -> pinned memory (async)."""
        c = self.codec
        B = locs.shape[0]
        l2, s2 = locs.reshape(B, -1), scales.reshape(B, -1)
        mm = np.stack([np.asarray(min_vs, np.int32).reshape(-1), np.asarray(max_vs, np.int32).reshape(-1)], -1)
        rows, off = c.laplace_cdf(l2, s2, mm, self._likelihood_bound)
        stage, done = runtime.to_host_async(rows, "cdf_rows%d" % slot)
        return stage, done, off, mm, l2.shape[1]

    @staticmethod
    def decode_finish(strings, stage, done, off, mm, E, slot: int, threads: int = 0):
        """Host half: -> pinned float32 torch tensor [B, E] of y_hat."""
        done.synchronize()
        return runtime.range_decode_rows_batch_f32(list(strings), E, stage.numpy(), off, mm, threads, out_tag="y_hat_dec%d" % slot)

    # ---- GPU-side coder (csrc/gpu_coder.cu): nothing per-element crosses PCIe ------------------------------------------
    def intervals_dev(self, ys, locs, scales, iv_out=None, want_likelihoods=False):
        """Quantise + per-cube range + per-element intervals, all on the device and asynchronous.
        -> (iv int32 [B,E], minmax int32 [B,2]) device tensors; ``iv_out`` = a [B,E] slice to fill in place."""
        c = self.codec
        B = ys.shape[0]
        y2, l2, s2 = ys.reshape(B, -1), locs.reshape(B, -1), scales.reshape(B, -1)
        y_hat, _, _, mm = c.laplace(y2, l2, s2, self._likelihood_bound, want_p=want_likelihoods, want_bits=want_likelihoods)
        if WIDEN_RANGES:
            c.widen_symbol_ranges(mm)
        iv = c.laplace_intervals(y_hat, l2, s2, mm, self._likelihood_bound, out=iv_out)
        return iv, mm

    def encode_dev(self, iv):
        """iv int32 [B,E] on the device -> (packed uint8, offsets int64 [B+1]) device tensors (asynchronous)."""
        return self.codec.gpu_range_encode(iv)

    def decode_rows_dev(self, locs, scales, min_vs, max_vs):
        """First half of ``decode_dev``: the per-element CDF rows of the cubes, built on the device from (loc, scale) and the
        header's symbol ranges (asynchronous on the current stream).  -> what ``decode_strings_dev`` needs."""
        c = self.codec
        B = locs.shape[0]
        l2, s2 = locs.reshape(B, -1), scales.reshape(B, -1)
        E = l2.shape[1]
        mm = np.stack([np.asarray(min_vs, np.int32).reshape(-1), np.asarray(max_vs, np.int32).reshape(-1)], -1)
        n_sym = mm[:, 1] - mm[:, 0] + 1
        if B and (n_sym.min() < 2 or n_sym.max() > 64):
            raise runtime._lib.PcgcError(runtime._lib.ERR_BAD_RANGE, "symbol range outside [2, 64] symbols")
        off = c.row_offsets(mm, E)
        hdr = np.concatenate([off, mm.reshape(-1).astype(np.int64)])             # one small H2D for both headers
        hdr_d = c.to_device(hdr)
        off_d, mm_d = hdr_d[:B + 1], hdr_d[B + 1:].to(torch.int32)
        rows = c.laplace_cdf_dev(l2, s2, mm_d, off_d, int(off[-1]), self._likelihood_bound)
        return rows, off_d, int(off[-1]), mm_d, (int(n_sym.max()) if B else 2), B, E

    def decode_strings_dev(self, packed, offsets, rows_pack):
        """Second half: strings b = packed[offsets[b]:offsets[b+1]] (device) read against the rows -> y_hat float32 [B,E]."""
        rows, off_d, total, mm_d, max_n, B, E = rows_pack
        return self.codec.gpu_range_decode(packed, offsets, rows, off_d, total, mm_d, max_n, B, E)

    def decode_dev(self, packed, offsets, locs, scales, min_vs, max_vs):
        """Strings b = packed[offsets[b]:offsets[b+1]] (device) -> y_hat float32 [B,E] on the device; CDF rows are built and
        consumed on the device (asynchronous on the current stream; headers come from the host)."""
        return self.decode_strings_dev(packed, offsets, self.decode_rows_dev(locs, scales, min_vs, max_vs))

    def decompress_cubes(self, strings, locs, scales, min_vs, max_vs, threads: int = 0):
        """-> torch float32 [B, E] on the device."""
        c = self.codec
        B = locs.shape[0]
        l2, s2 = locs.reshape(B, -1), scales.reshape(B, -1)
        E = l2.shape[1]
        if runtime.coder_mode() == "gpu" and B > 0:
            packed, offsets = c.upload_strings(list(strings))
            y_hat = self.decode_dev(packed, offsets, l2, s2, min_vs, max_vs)
            c.synchronize()
            return y_hat
        mm = np.stack([np.asarray(min_vs, np.int32).reshape(-1), np.asarray(max_vs, np.int32).reshape(-1)], -1)
        rows, off = c.laplace_cdf(l2, s2, mm, self._likelihood_bound)
        y_hat = runtime.range_decode_rows_batch_f32(list(strings), E, runtime.to_host(rows, "cdf_rows"), off, mm, threads)
        return c.to_device(y_hat)
