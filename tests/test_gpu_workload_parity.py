"""Oracle comparisons on cubes of the BASELINE workloads themselves (not only on small synthetic surface cubes):
config 1 (vox10, hyper mode), config 3 (sparse vox12 cubes), config 2 (model_simple, factorized) and the batch sizes of config 4.
Same tolerances as tests/test_gpu_parity.py (north_star): quantised latents identical on >= 99.9 % of the elements with every
mismatch on a rounding boundary, likelihoods / bits within 1e-3 relative, voxel set exact given identical latents."""
import numpy as np
import pytest
import torch

from oracle import entropy, nets, topk
from pcgcv1_b200 import runtime, synthetic, transform, weights as W
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import model_simple, model_voxception

pytestmark = pytest.mark.gpu

LIK_RTOL, LIK_ATOL = 1e-3, 3e-7


def _agreement(a_gpu, a_ref, a_ref64=None):
    qa, qb = np.rint(a_gpu), np.rint(a_ref)
    bad = qa != qb
    if bad.any():
        ref = a_ref64 if a_ref64 is not None else a_ref
        dist = np.abs(np.abs(ref[bad] - np.floor(ref[bad])) - 0.5)
        assert dist.max() < 5e-3, "a rounding mismatch is not a boundary case: %g" % dist.max()
    return 1.0 - bad.mean(), int(bad.sum())


def _sample(name, n):
    """n cubes spread over the whole cloud (dense and sparse ones alike)."""
    cubes, pos, nums = synthetic.workload(name, seed=0)
    idx = np.unique(np.linspace(0, len(cubes) - 1, n).astype(np.int64))
    return cubes[idx], nums[idx], len(cubes)


def _hyper_path_vs_oracle(codec, cubes, nums):
    w = W.synthetic_weights("voxception")
    x = cubes.astype(np.float32)
    part = lambda k: W.net_weights(w, k)
    y_ref = nets.run_net("voxception", "analysis", x, part("analysis_transform"))
    y_ref64 = nets.run_net("voxception", "analysis", x, part("analysis_transform"), dtype=torch.float64)
    y_gpu = codec.analysis(codec.to_device(cubes))
    frac, nbad = _agreement(y_gpu.cpu().numpy(), y_ref, y_ref64)
    print("y: %d cubes, quantised latents identical on %.6f (%d mismatches, all on rounding boundaries)" % (len(cubes), frac, nbad))
    assert frac >= 0.999
    assert np.abs(y_gpu.cpu().numpy() - y_ref).max() < 3e-4 * max(1.0, np.abs(y_ref).max())
    # hyper latents
    z_ref = nets.run_net("voxception", "hyper_encoder", y_gpu.cpu().numpy(), part("hyper_encoder"))
    z_gpu = codec.hyper_encode(y_gpu)
    fz, nz = _agreement(z_gpu.cpu().numpy(), z_ref)
    assert fz >= 0.999
    z_hat = torch.round(z_gpu)
    loc_ref, scale_ref = nets.run_net("voxception", "hyper_decoder", z_hat.cpu().numpy(), part("hyper_decoder"))
    scale_ref = np.maximum(scale_ref, np.float32(1e-9))
    loc, scale = codec.hyper_decode(z_hat, 1e-9)
    assert np.abs(loc.cpu().numpy() - loc_ref).max() < 3e-4 * max(1.0, np.abs(loc_ref).max())
    assert np.abs(scale.cpu().numpy() - scale_ref).max() < 3e-4 * max(1.0, np.abs(scale_ref).max())
    # likelihoods, bits, per-cube ranges on the GPU's own (y, loc, scale)
    B = len(cubes)
    y_hat, p, bits, mm = codec.laplace(y_gpu.reshape(B, -1), loc.reshape(B, -1), scale.reshape(B, -1))
    ref_hat, ref_p = entropy.SymmetricConditionalOracle()(y_gpu.cpu().numpy().reshape(B, -1), loc.cpu().numpy().reshape(B, -1),
                                                           scale.cpu().numpy().reshape(B, -1))
    assert np.array_equal(y_hat.cpu().numpy(), ref_hat)
    np.testing.assert_allclose(p.cpu().numpy(), ref_p, rtol=LIK_RTOL, atol=LIK_ATOL)
    for b in range(B):
        rb = entropy.estimated_bits(ref_p[b])
        assert abs(float(bits[b]) - rb) <= 1e-3 * rb
        assert (int(mm[b, 0]), int(mm[b, 1])) == (int(ref_hat[b].min()), int(ref_hat[b].max()))
    # full codec: stream -> identical reconstruction -> voxel set == oracle top-k on the same logits
    out = transform.compress_hyper(cubes, model_voxception, "", decompress=True)
    xs = transform.decompress_hyper(*[o.numpy() for o in out[:8]], model_voxception, "")
    assert torch.equal(xs.tensor, out[8].tensor)
    logits_ref = nets.run_net("voxception", "synthesis", y_hat.cpu().numpy().reshape(y_ref.shape), part("synthesis_transform"))
    assert np.abs(xs.numpy() - logits_ref).max() < 1e-3 * max(1.0, np.abs(logits_ref).max())
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec)
    assert np.array_equal(mask, topk.select_voxels(xs.numpy(), nums, 1.0))
    return out


def test_config1_vox10_sample_vs_oracle(codec):
    cubes, nums, total = _sample("vox10", 16)
    assert total >= 150 and len(cubes) == 16
    _hyper_path_vs_oracle(codec, cubes, nums)


def test_config3_vox12_sample_vs_oracle(codec):
    cubes, nums, total = _sample("vox12", 12)
    assert total >= 2000                      # thousands of lightly filled cubes
    assert nums.max() < 2000                  # sparse: nothing like the ~4k points of a vox10 surface cube
    _hyper_path_vs_oracle(codec, cubes, nums)


def test_config2_simple_factorized_sample_vs_oracle(codec_simple):
    cubes, nums, _ = _sample("vox10", 16)
    w = W.synthetic_weights("simple")
    x = cubes.astype(np.float32)
    y_ref = nets.run_net("simple", "analysis", x, W.net_weights(w, "analysis_transform"))
    y_gpu = codec_simple.analysis(codec_simple.to_device(cubes)).cpu().numpy()
    frac, nbad = _agreement(y_gpu, y_ref)
    print("simple y: identical on %.6f (%d mismatches)" % (frac, nbad))
    assert frac >= 0.999
    strings, min_v, max_v, shape = transform.compress_factorized(cubes, model_simple, "")
    xs = transform.decompress_factorized(strings.numpy(), min_v.numpy(), max_v.numpy(), shape.numpy(), model_simple, "")
    logits_ref = nets.run_net("simple", "synthesis", np.rint(y_gpu), W.net_weights(w, "synthesis_transform"))
    assert np.abs(xs.numpy() - logits_ref).max() < 1e-3 * max(1.0, np.abs(logits_ref).max())
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec_simple)
    assert np.array_equal(mask, topk.select_voxels(xs.numpy(), nums, 1.0))
    # factorized likelihoods / bits of y against the oracle's EntropyBottleneck (32 channels)
    eb = entropy.EntropyBottleneckOracle(W.net_weights(w, "estimator"))
    slot = codec_simple.bottleneck_slot(32)
    yd = codec_simple.to_device(y_gpu)
    x_hat, p, bits, mm = codec_simple.factorized(slot, yd)
    ref_hat, ref_p = eb(y_gpu)
    assert np.array_equal(x_hat.cpu().numpy(), ref_hat)
    np.testing.assert_allclose(p.cpu().numpy(), ref_p, rtol=LIK_RTOL, atol=LIK_ATOL)
    rb = entropy.estimated_bits(ref_p)
    assert abs(float(bits[0]) - rb) <= 1e-3 * rb


def test_config4_batch_512_is_batch_invariant(codec):
    """The largest batch of the config-4 sweep: every cube's latents and logits equal what the same cube gives alone
    (tile shapes never depend on the batch; README.md:111-114 is the failure mode this excludes)."""
    base, nums = synthetic.surface_cubes(8, seed=3)
    reps = 64
    big = np.concatenate([base] * reps)                                   # 512 cubes
    assert len(big) == 512
    xb = codec.to_device(big)
    y_big = codec.analysis(xb)
    y_one = codec.analysis(codec.to_device(base))
    for r in (0, 17, 63):
        assert torch.equal(y_big[8 * r:8 * r + 8], y_one)
    g_big = codec.synthesis(torch.round(y_big))
    g_one = codec.synthesis(torch.round(y_one))
    for r in (0, 31, 63):
        assert torch.equal(g_big[8 * r:8 * r + 8], g_one)
    del y_big, g_big, xb
    torch.cuda.empty_cache()
