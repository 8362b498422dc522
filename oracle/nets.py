"""Oracle restatement of the reference's conv transforms (test infrastructure only).

Follows ``models/model_voxception.py:11-308`` and ``models/model_simple.py:12-95`` layer by
layer.  Data layout matches the reference (channels-last NDHWC); weights use the Keras
layouts: Conv3D kernel ``[kd,kh,kw,Cin,Cout]``, Conv3DTranspose kernel ``[kd,kh,kw,Cout,Cin]``,
bias ``[Cout]``.  torch-CPU ``conv3d``/``conv_transpose3d`` do the arithmetic; the TF/Keras
semantics that live outside the reference tree are restated here:

* SAME padding for stride s: ``total = (ceil(n/s)-1)*s + k - n``, ``before = total//2``,
  ``after = total - before``  (k3 s2 -> (0,1); k5 s2 -> (1,2); k9 s2 -> (3,4); s1 -> symmetric).
* ``Conv3DTranspose(padding='same', strides=2)``: output ``2n``, equal to the full transposed
  convolution (kernel un-flipped) cropped to ``[before : before+2n]`` per axis.
* Keras layer order: conv -> bias add -> activation.

``dtype=torch.float32`` gives the "O32" oracle (what a fp32 TF-CPU run would produce up to
summation order), ``torch.float64`` gives "O64" (ground truth for guard-band decisions).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Weights = Dict[str, np.ndarray]


def _same_pads(n: int, k: int, s: int) -> Tuple[int, int]:
    total = max((math.ceil(n / s) - 1) * s + k - n, 0)
    before = total // 2
    return before, total - before


def _t(a, dtype) -> torch.Tensor:
    if isinstance(a, torch.Tensor):                      # training oracle: weights are leaf tensors that require grad
        return a.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def conv3d_same(x: torch.Tensor, w: Weights, name: str, stride: int = 1, relu: bool = False) -> torch.Tensor:
    """tf.keras.layers.Conv3D(padding='same') on an NDHWC tensor."""
    kern = w[name + "/kernel"]                      # [kd,kh,kw,Cin,Cout]
    k = kern.shape[0]
    wt = _t(kern, x.dtype).permute(4, 3, 0, 1, 2).contiguous()   # [Cout,Cin,kd,kh,kw]
    bias = _t(w[name + "/bias"], x.dtype) if (name + "/bias") in w else None
    xc = x.permute(0, 4, 1, 2, 3)
    pads = []
    for n in (xc.shape[4], xc.shape[3], xc.shape[2]):            # F.pad wants last dim first
        b, a = _same_pads(n, k, stride)
        pads += [b, a]
    xc = F.pad(xc, pads)
    y = F.conv3d(xc, wt, bias, stride=stride)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def conv3d_transpose_same(x: torch.Tensor, w: Weights, name: str, stride: int = 2, relu: bool = False) -> torch.Tensor:
    """tf.keras.layers.Conv3DTranspose(padding='same', strides=2) on an NDHWC tensor."""
    kern = w[name + "/kernel"]                      # [kd,kh,kw,Cout,Cin]
    k = kern.shape[0]
    wt = _t(kern, x.dtype).permute(4, 3, 0, 1, 2).contiguous()   # torch wants [Cin,Cout,kd,kh,kw]
    bias = _t(w[name + "/bias"], x.dtype) if (name + "/bias") in w else None
    xc = x.permute(0, 4, 1, 2, 3)
    full = F.conv_transpose3d(xc, wt, None, stride=stride)       # length (n-1)*s + k
    n = xc.shape[2]
    pb, _ = _same_pads(n * stride, k, stride)
    y = full[:, :, pb:pb + stride * xc.shape[2], pb:pb + stride * xc.shape[3], pb:pb + stride * xc.shape[4]]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def vrn_block(x: torch.Tensor, w: Weights, name: str) -> torch.Tensor:
    """_VoxceptionResNet.call, models/model_voxception.py:56-68."""
    t11 = conv3d_same(x, w, name + "_conv1_1", relu=True)
    t12 = conv3d_same(t11, w, name + "_conv1_2", relu=True)
    t21 = conv3d_same(x, w, name + "_conv2_1", relu=True)
    t22 = conv3d_same(t21, w, name + "_conv2_2", relu=True)
    t23 = conv3d_same(t22, w, name + "_conv2_3", relu=True)
    return torch.relu(x + torch.cat([t12, t23], dim=-1))


def analysis_voxception(x: torch.Tensor, w: Weights) -> torch.Tensor:
    """AnalysisTransform.call, models/model_voxception.py:125-144."""
    f = conv3d_same(x, w, "conv_in", relu=True)
    for i in (1, 2, 3):
        f = vrn_block(f, w, "vrn1_%d" % i)
    f = conv3d_same(f, w, "down_1", stride=2, relu=True)
    for i in (1, 2, 3):
        f = vrn_block(f, w, "vrn2_%d" % i)
    f = conv3d_same(f, w, "down_2", stride=2, relu=True)
    for i in (1, 2, 3):
        f = vrn_block(f, w, "vrn3_%d" % i)
    return conv3d_same(f, w, "conv_out")


def synthesis_voxception(y: torch.Tensor, w: Weights) -> torch.Tensor:
    """SynthesisTransform.call, models/model_voxception.py:195-214."""
    f = conv3d_same(y, w, "deconv_in", relu=True)
    for i in (1, 2, 3):
        f = vrn_block(f, w, "dvrn1_%d" % i)
    f = conv3d_transpose_same(f, w, "up_1", relu=True)
    for i in (1, 2, 3):
        f = vrn_block(f, w, "dvrn2_%d" % i)
    f = conv3d_transpose_same(f, w, "up_2", relu=True)
    for i in (1, 2, 3):
        f = vrn_block(f, w, "dvrn3_%d" % i)
    return conv3d_same(f, w, "deconv_out")


def hyper_encoder(y: torch.Tensor, w: Weights) -> torch.Tensor:
    """HyperEncoder.call, models/model_voxception.py:246-252."""
    f = conv3d_same(y, w, "conv1", relu=True)
    f = conv3d_same(f, w, "conv2", stride=2, relu=True)
    return conv3d_same(f, w, "conv3")


def hyper_decoder(z: torch.Tensor, w: Weights) -> Tuple[torch.Tensor, torch.Tensor]:
    """HyperDecoder.call, models/model_voxception.py:299-308: returns (loc, abs(scale))."""
    f = conv3d_same(z, w, "deconv1", relu=True)
    f = conv3d_transpose_same(f, w, "deconv2", relu=True)
    f = conv3d_same(f, w, "deconv3", relu=True)
    loc = conv3d_same(f, w, "deconv4_1")
    scale = conv3d_same(f, w, "deconv4_2")
    return loc, torch.abs(scale)


def analysis_simple(x: torch.Tensor, w: Weights) -> torch.Tensor:
    """model_simple.AnalysisTransform.call, models/model_simple.py:45-51."""
    f = conv3d_same(x, w, "conv_1", stride=2, relu=True)
    f = conv3d_same(f, w, "conv_2", stride=2, relu=True)
    return conv3d_same(f, w, "conv_3", stride=2)


def synthesis_simple(y: torch.Tensor, w: Weights) -> torch.Tensor:
    """model_simple.SynthesisTransform.call, models/model_simple.py:89-95."""
    f = conv3d_transpose_same(y, w, "deconv_1", relu=True)
    f = conv3d_transpose_same(f, w, "deconv_2", relu=True)
    return conv3d_transpose_same(f, w, "deconv_3")


NETS = {
    ("voxception", "analysis"): analysis_voxception,
    ("voxception", "synthesis"): synthesis_voxception,
    ("voxception", "hyper_encoder"): hyper_encoder,
    ("voxception", "hyper_decoder"): hyper_decoder,
    ("simple", "analysis"): analysis_simple,
    ("simple", "synthesis"): synthesis_simple,
}


def run_net(model: str, net: str, x: np.ndarray, w: Weights, dtype=torch.float32, per_cube: bool = False):
    """Run one net on a NumPy NDHWC batch.  ``per_cube=True`` drives it the way the
    reference does: one cube per call (tf.map_fn(parallel_iterations=1), transform.py:48,122)."""
    fn = NETS[(model, net)]
    xt = _t(np.asarray(x), dtype)
    with torch.no_grad():
        if per_cube:
            outs = [fn(xt[i:i + 1], w) for i in range(xt.shape[0])]
            if isinstance(outs[0], tuple):
                out = tuple(torch.cat([o[j] for o in outs]) for j in range(len(outs[0])))
            else:
                out = torch.cat(outs)
        else:
            out = fn(xt, w)
    if isinstance(out, tuple):
        return tuple(o.numpy() for o in out)
    return out.numpy()
