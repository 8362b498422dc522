"""Layer tables of the reference's transforms (host side; no arithmetic here).

One ``Layer`` per Keras layer, in call order, with the Keras ``name=`` so that a weight file
keyed like the reference's checkpoint (``<net>/<layer>/kernel|bias``) loads directly.
Sources: ``models/model_voxception.py:21-54,83-122,153-192,224-244,263-297`` and
``models/model_simple.py:21-42,58-86``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class Layer:
    name: str
    cin: int
    cout: int
    k: int = 3
    stride: int = 1
    transposed: bool = False
    bias: bool = True
    relu: bool = True


def _vrn(name: str, c: int) -> List[Layer]:
    # _VoxceptionResNet, model_voxception.py:21-54
    return [
        Layer(name + "_conv1_1", c, c // 4, 3),
        Layer(name + "_conv1_2", c // 4, c // 2, 3),
        Layer(name + "_conv2_1", c, c // 4, 1),
        Layer(name + "_conv2_2", c // 4, c // 4, 3),
        Layer(name + "_conv2_3", c // 4, c // 2, 1),
    ]


def voxception_analysis() -> List[Layer]:
    ls = [Layer("conv_in", 1, 16)]
    for i in (1, 2, 3):
        ls += _vrn("vrn1_%d" % i, 16)
    ls.append(Layer("down_1", 16, 32, 3, 2, bias=False))
    for i in (1, 2, 3):
        ls += _vrn("vrn2_%d" % i, 32)
    ls.append(Layer("down_2", 32, 64, 3, 2, bias=False))
    for i in (1, 2, 3):
        ls += _vrn("vrn3_%d" % i, 64)
    ls.append(Layer("conv_out", 64, 16, relu=False))
    return ls


def voxception_synthesis() -> List[Layer]:
    ls = [Layer("deconv_in", 16, 64)]
    for i in (1, 2, 3):
        ls += _vrn("dvrn1_%d" % i, 64)
    ls.append(Layer("up_1", 64, 32, 3, 2, transposed=True))
    for i in (1, 2, 3):
        ls += _vrn("dvrn2_%d" % i, 32)
    ls.append(Layer("up_2", 32, 16, 3, 2, transposed=True))
    for i in (1, 2, 3):
        ls += _vrn("dvrn3_%d" % i, 16)
    ls.append(Layer("deconv_out", 16, 1, relu=False))
    return ls


def hyper_encoder() -> List[Layer]:
    return [Layer("conv1", 16, 16), Layer("conv2", 16, 16, 3, 2), Layer("conv3", 16, 8, relu=False)]


def hyper_decoder() -> List[Layer]:
    return [
        Layer("deconv1", 8, 16),
        Layer("deconv2", 16, 16, 3, 2, transposed=True),
        Layer("deconv3", 16, 32),
        Layer("deconv4_1", 32, 16, relu=False),
        Layer("deconv4_2", 32, 16, relu=False),
    ]


def simple_analysis() -> List[Layer]:
    return [
        Layer("conv_1", 1, 32, 9, 2),
        Layer("conv_2", 32, 32, 5, 2),
        Layer("conv_3", 32, 32, 5, 2, bias=False, relu=False),
    ]


def simple_synthesis() -> List[Layer]:
    return [
        Layer("deconv_1", 32, 32, 5, 2, transposed=True),
        Layer("deconv_2", 32, 32, 5, 2, transposed=True),
        Layer("deconv_3", 32, 1, 9, 2, transposed=True, relu=False),
    ]


# checkpoint top-level keys (transform.py:107-111) -> layer table
NETS: Dict[Tuple[str, str], List[Layer]] = {
    ("voxception", "analysis_transform"): voxception_analysis(),
    ("voxception", "synthesis_transform"): voxception_synthesis(),
    ("voxception", "hyper_encoder"): hyper_encoder(),
    ("voxception", "hyper_decoder"): hyper_decoder(),
    ("simple", "analysis_transform"): simple_analysis(),
    ("simple", "synthesis_transform"): simple_synthesis(),
}

LATENT_CHANNELS = {"voxception": 16, "simple": 32}      # channels of y
LATENT_DOWN = {"voxception": 4, "simple": 8}            # 64 -> 16 / 8
HYPER_CHANNELS = 8                                      # channels of z (voxception only)


def kernel_shape(l: Layer) -> Tuple[int, int, int, int, int]:
    """Keras layout: Conv3D [k,k,k,Cin,Cout]; Conv3DTranspose [k,k,k,Cout,Cin]."""
    return (l.k, l.k, l.k, l.cout, l.cin) if l.transposed else (l.k, l.k, l.k, l.cin, l.cout)


def macs_per_cube(layers: List[Layer], in_size: int) -> int:
    """Dense MACs per cube, as counted in SURVEY.md appendix A."""
    n, total = in_size, 0
    vrn_in = None
    for l in layers:
        if l.name.endswith("_conv1_1"):
            vrn_in = n
        if l.transposed:
            total += n ** 3 * l.k ** 3 * l.cin * l.cout
            n *= l.stride
        else:
            n_out = n // l.stride
            total += n_out ** 3 * l.k ** 3 * l.cin * l.cout
            n = n_out
    return total
