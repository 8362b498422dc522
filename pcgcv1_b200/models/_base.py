"""Shared base of the transform classes (host side)."""
from __future__ import annotations

from .. import runtime


class _Transform:
    """Zero-argument constructor + ``__call__`` on an ``[N,D,H,W,C]`` batch, like the reference's
    Keras models.  Weights come from a ``runtime.Codec``: the one ``transform.py`` binds, or
    the cached default codec (seeded synthetic weights, ``ckpt_dir == ''``)."""

    MODEL = "voxception"

    def __init__(self, codec=None):
        self._codec = codec

    def bind(self, codec):
        self._codec = codec
        return self

    @property
    def codec(self):
        if self._codec is None:
            self._codec = runtime.get_codec(self.MODEL, "")
        return self._codec
