"""Drop-in for ``models/model_simple.py`` (model_simple.py:12-95): the shallow 9^3/5^3/5^3 stride-2
autoencoder used in factorized mode."""
from __future__ import annotations

from .. import runtime
from ._base import _Transform

MODEL_NAME = "simple"


class AnalysisTransform(_Transform):
    """x [N,64,64,64,1] -> y [N,8,8,8,32]  (model_simple.py:45-51)."""
    MODEL = "simple"

    def __call__(self, x):
        c = self.codec
        return runtime.DeviceResult(c.analysis(c.to_device(x)))


class SynthesisTransform(_Transform):
    """y [N,8,8,8,32] -> logits [N,64,64,64,1]  (model_simple.py:89-95)."""
    MODEL = "simple"

    def __call__(self, y):
        c = self.codec
        import torch
        return runtime.DeviceResult(c.synthesis(c.to_device(y, torch.float32)))
