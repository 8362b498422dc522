"""Drop-in for ``models/model_voxception.py``: AnalysisTransform / SynthesisTransform / HyperEncoder /
HyperDecoder with the reference's constructor and call signatures (model_voxception.py:71-308).
The layer graphs live in libpcgc_b200.so (csrc/api.cu); these classes only route buffers."""
from __future__ import annotations

from .. import runtime
from ._base import _Transform

MODEL_NAME = "voxception"


class AnalysisTransform(_Transform):
    """x [N,64,64,64,1] -> y [N,16,16,16,16]  (model_voxception.py:125-144)."""

    def __call__(self, x):
        c = self.codec
        return runtime.DeviceResult(c.analysis(c.to_device(x)))


class SynthesisTransform(_Transform):
    """y [N,16,16,16,16] -> occupancy logits [N,64,64,64,1]  (model_voxception.py:195-214)."""

    def __call__(self, y):
        c = self.codec
        import torch
        return runtime.DeviceResult(c.synthesis(c.to_device(y, torch.float32)))


class HyperEncoder(_Transform):
    """y [N,16,16,16,16] -> z [N,8,8,8,8]  (model_voxception.py:246-252)."""

    def __call__(self, y):
        c = self.codec
        import torch
        return runtime.DeviceResult(c.hyper_encode(c.to_device(y, torch.float32)))


class HyperDecoder(_Transform):
    """z [N,8,8,8,8] -> (loc, abs(scale)) each [N,16,16,16,16]  (model_voxception.py:299-308).
    The caller applies ``max(scale, 1e-9)`` (transform.py:146); here scale_floor=0 keeps |scale|."""

    def __call__(self, z, scale_floor: float = 0.0):
        c = self.codec
        import torch
        loc, scale = c.hyper_decode(c.to_device(z, torch.float32), scale_floor)
        return runtime.DeviceResult(loc), runtime.DeviceResult(scale)
