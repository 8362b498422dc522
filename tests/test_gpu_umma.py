"""The tcgen05 (UMMA) conv engine against a plain PyTorch FP32 conv of the same op, layer shape by layer
shape, through the C ABI test hook pcgc_debug_conv3_umma.  Tolerance: the split-bf16 scheme drops only the
x_lo*w_lo term (~2^-17 relative per product); with FP32 accumulation the result must agree with an FP32
conv to ~2e-5 of the output scale."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# (grid n, Cin, Cout): every distinct stride-1 3x3x3 shape the engine serves (fused VRN shapes included)
# (grid n, Cin, Cout, WT): WT > 1 = the y-banded form (each M row produces WT output lines)
SHAPES = [(64, 16, 8, 1), (64, 8, 12, 1), (64, 16, 1, 1), (32, 32, 16, 1), (32, 16, 24, 1), (16, 64, 32, 1), (16, 32, 48, 1), (16, 64, 16, 1),
          (16, 16, 64, 1), (16, 16, 16, 1),
          (64, 16, 8, 2), (64, 8, 12, 2), (64, 16, 1, 4), (32, 32, 16, 2), (32, 16, 24, 2)]


@pytest.mark.parametrize("n,cin,cout,wt", SHAPES)
def test_umma_conv_vs_torch_fp32(codec, n, cin, cout, wt):
    g = torch.Generator(device="cpu").manual_seed(n * 1000 + cin * 10 + cout)
    B = 2
    x = torch.randn(B, n, n, n, cin, generator=g).relu_()            # post-ReLU like real activations
    x[:, : n // 2] *= 37.0                                             # mixed magnitudes
    w = torch.randn(3, 3, 3, cin, cout, generator=g) * (2.0 / (27 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    xd = x.to(codec.dev)
    out = torch.empty(B, n, n, n, cout, device=codec.dev)
    wh, bh = np.ascontiguousarray(w.numpy()), np.ascontiguousarray(b.numpy())
    codec._stream()
    rc = codec.lib.pcgc_debug_conv3_umma(codec.ctx, xd.data_ptr(), n, cin, cout, wh.ctypes.data, bh.ctypes.data, 1, B, wt, out.data_ptr())
    codec._check(rc)
    ref = torch.nn.functional.conv3d(xd.permute(0, 4, 1, 2, 3).double(), w.to(codec.dev).permute(4, 3, 0, 1, 2).double(),
                                     b.to(codec.dev).double(), padding=1).relu_().permute(0, 2, 3, 4, 1)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print("n=%d cin=%d cout=%d wt=%d: max abs err %.3g (scale %.3g, rel %.2g)" % (n, cin, cout, wt, err, scale, err / scale))
    assert err <= 3e-5 * scale
    # deterministic
    out2 = torch.empty_like(out)
    codec._check(codec.lib.pcgc_debug_conv3_umma(codec.ctx, xd.data_ptr(), n, cin, cout, wh.ctypes.data, bh.ctypes.data, 1, B, wt, out2.data_ptr()))
    assert torch.equal(out, out2)


def test_engines_agree_on_the_transforms(codec):
    """tcgen05 engine (fused VRN kernels, split-bf16) vs the exact-FP32 CUDA-core engine on the same cubes."""
    from pcgcv1_b200 import _lib, synthetic
    cubes, _ = synthetic.surface_cubes(3, seed=21)
    x = codec.to_device(cubes)
    try:
        codec.set_engine(_lib.ENGINE_FFMA)
        y_f = codec.analysis(x)
        g_f = codec.synthesis(torch.round(y_f))
        codec.set_engine(_lib.ENGINE_AUTO)
        y_u = codec.analysis(x)
        g_u = codec.synthesis(torch.round(y_f))
        codec.synchronize()
    finally:
        codec.set_engine(_lib.ENGINE_AUTO)
    ey = (y_u - y_f).abs().max().item()
    eg = (g_u - g_f).abs().max().item()
    agree = (torch.round(y_u) == torch.round(y_f)).float().mean().item()
    print("analysis |d|max %.3g (|y|max %.3g), synthesis |d|max %.3g (|logit|max %.3g), rounding agreement %.6f"
          % (ey, y_f.abs().max().item(), eg, g_f.abs().max().item(), agree))
    assert ey < 1e-3 * max(1.0, y_f.abs().max().item())
    assert eg < 1e-3 * max(1.0, g_f.abs().max().item())
    assert agree >= 0.999
    # bit-reproducible and batch-invariant on the tcgen05 engine too
    assert torch.equal(y_u, codec.analysis(x))
    assert torch.equal(y_u[2:3], codec.analysis(x[2:3]))


def test_simple_model_engines_agree(codec_simple):
    """model_simple (model_simple.py:21-42,58-86): window-GEMM tcgen05 kernel (umma_win.cu; stride-2 layers as stride-1 windows
    over space-to-depth channels / output-parity columns, split-bf16) vs the exact-FP32 CUDA-core engine, layer chain by layer
    chain, on non-binary float cubes too (the occupancy grid is converted with a hi and a lo plane), different batch sizes."""
    from pcgcv1_b200 import _lib, synthetic
    codec = codec_simple
    cubes, _ = synthetic.surface_cubes(5, seed=33)
    rng = np.random.default_rng(5)
    soft = (cubes.astype(np.float32) * rng.uniform(0.25, 1.0, size=cubes.shape).astype(np.float32))      # not representable in bf16
    for x_host in (cubes, soft):
        x = codec.to_device(x_host)
        try:
            codec.set_engine(_lib.ENGINE_FFMA)
            y_f = codec.analysis(x)
            g_f = codec.synthesis(torch.round(y_f))
            codec.set_engine(_lib.ENGINE_AUTO)
            y_u = codec.analysis(x)
            g_u = codec.synthesis(torch.round(y_f))
            codec.synchronize()
        finally:
            codec.set_engine(_lib.ENGINE_AUTO)
        ey = (y_u - y_f).abs().max().item()
        eg = (g_u - g_f).abs().max().item()
        agree = (torch.round(y_u) == torch.round(y_f)).float().mean().item()
        print("simple: analysis |d|max %.3g (|y|max %.3g), synthesis |d|max %.3g (|logit|max %.3g), rounding agreement %.6f"
              % (ey, y_f.abs().max().item(), eg, g_f.abs().max().item(), agree))
        assert ey < 1e-4 * max(1.0, y_f.abs().max().item())
        assert eg < 1e-4 * max(1.0, g_f.abs().max().item())
        assert agree >= 0.999
        # bit-reproducible and batch-invariant
        assert torch.equal(y_u, codec.analysis(x))
        assert torch.equal(y_u[3:4], codec.analysis(x[3:4]))
        assert torch.equal(g_u[1:3], codec.synthesis(torch.round(y_f)[1:3]))


def test_simple_model_vs_fp64_reference(codec_simple):
    """model_simple's analysis and synthesis chains against float64 torch convs with the TF SAME / Conv3DTranspose rules
    (oracle.nets, pinned by tests/golden/golden_nets.npz): tolerance 2e-5 of the output scale, which a wrong tap, parity class or
    crop offset in any layer misses by orders of magnitude (the FP32 CUDA-core engine sits at ~1e-6, split-bf16 at ~5e-6)."""
    from oracle import nets
    from pcgcv1_b200 import synthetic, weights as W
    w = W.synthetic_weights("simple")
    cubes, _ = synthetic.surface_cubes(2, seed=44)
    y = codec_simple.analysis(codec_simple.to_device(cubes)).cpu().numpy().astype(np.float64)
    ref = nets.run_net("simple", "analysis", cubes.astype(np.float64), W.net_weights(w, "analysis_transform"), dtype=torch.float64)
    assert np.abs(y - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    yq = np.rint(ref)
    g = codec_simple.synthesis(codec_simple.to_device(yq.astype(np.float32))).cpu().numpy().astype(np.float64)
    ref_g = nets.run_net("simple", "synthesis", yq, W.net_weights(w, "synthesis_transform"), dtype=torch.float64)
    assert np.abs(g - ref_g).max() < 2e-5 * max(1.0, np.abs(ref_g).max())
