"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the golden vectors.
Run on the GPU box: ``python -m pytest tests -m gpu``.

Tolerances (BASELINE.json north_star): quantised latents bit-exact on >= 99.9 % of elements (every
mismatch must sit on a rounding boundary); likelihoods / estimated bits within 1e-3 relative
(likelihoods get an absolute floor of 3e-7: the reference's own formula cancels catastrophically in
float32 for tail symbols); decoded voxel set exact given identical latents.
"""
import numpy as np
import pytest
import torch

from oracle import coder, entropy, nets, topk
from pcgcv1_b200 import _lib, runtime, synthetic, transform, weights as W
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import conditional_entropy_model, entropy_model, model_simple, model_voxception

pytestmark = pytest.mark.gpu

LIK_RTOL, LIK_ATOL = 1e-3, 3e-7


def _round_agreement(a_gpu, a_ref, a_ref64=None):
    """fraction of elements whose rounding agrees; mismatches must be rounding-boundary cases."""
    qa, qb = np.rint(a_gpu), np.rint(a_ref)
    bad = qa != qb
    frac = 1.0 - bad.mean()
    if bad.any():
        ref = a_ref64 if a_ref64 is not None else a_ref
        dist = np.abs(np.abs(ref[bad] - np.floor(ref[bad])) - 0.5)
        assert dist.max() < 5e-3, "a rounding mismatch is not a boundary case: %g" % dist.max()
    return frac, int(bad.sum())


@pytest.fixture(scope="module")
def cubes():
    c, n = synthetic.surface_cubes(3, seed=11)
    return c, n


@pytest.fixture(scope="module")
def oracle_latents(cubes):
    """O32 oracle of the whole hyper path for the 3 test cubes (a few seconds of CPU)."""
    c, _ = cubes
    w = W.synthetic_weights("voxception")
    x = c.astype(np.float32)
    y = nets.run_net("voxception", "analysis", x, W.net_weights(w, "analysis_transform"))
    y64 = nets.run_net("voxception", "analysis", x, W.net_weights(w, "analysis_transform"), dtype=torch.float64)
    z = nets.run_net("voxception", "hyper_encoder", y, W.net_weights(w, "hyper_encoder"))
    loc, scale = nets.run_net("voxception", "hyper_decoder", np.rint(z), W.net_weights(w, "hyper_decoder"))
    scale = np.maximum(scale, np.float32(1e-9))
    logits = nets.run_net("voxception", "synthesis", np.rint(y), W.net_weights(w, "synthesis_transform"))
    return dict(w=w, y=y, y64=y64, z=z, loc=loc, scale=scale, logits=logits)


# ------------------------------------------------------------------------------- transforms
def test_analysis_parity(codec, cubes, oracle_latents):
    y = codec.analysis(codec.to_device(cubes[0])).cpu().numpy()
    ref = oracle_latents["y"]
    err = np.abs(y - ref).max()
    assert err < 2e-4 * max(1.0, np.abs(ref).max()), err
    frac, nbad = _round_agreement(y, ref, oracle_latents["y64"])
    print("analysis: max abs err %.3g, rounding agreement %.6f (%d mismatches)" % (err, frac, nbad))
    assert frac >= 0.999


def test_analysis_accepts_float_inputs(codec, cubes):
    x8 = codec.to_device(cubes[0][:1])
    a = codec.analysis(x8)
    b = codec.analysis(x8.to(torch.float32))
    c = codec.analysis(x8.to(torch.float64))
    assert torch.equal(a, b) and torch.equal(a, c)


def test_hyper_and_synthesis_parity(codec, oracle_latents):
    o = oracle_latents
    z = codec.hyper_encode(codec.to_device(o["y"])).cpu().numpy()
    assert np.abs(z - o["z"]).max() < 2e-4 * max(1.0, np.abs(o["z"]).max())
    assert _round_agreement(z, o["z"])[0] >= 0.999
    loc, scale = codec.hyper_decode(codec.to_device(np.rint(o["z"])), 1e-9)
    assert np.abs(loc.cpu().numpy() - o["loc"]).max() < 1e-4 * max(1.0, np.abs(o["loc"]).max())
    assert np.abs(scale.cpu().numpy() - o["scale"]).max() < 1e-4 * max(1.0, np.abs(o["scale"]).max())
    logits = codec.synthesis(codec.to_device(np.rint(o["y"]))).cpu().numpy()
    assert np.abs(logits - o["logits"]).max() < 2e-4 * max(1.0, np.abs(o["logits"]).max())


def test_transform_classes_mirror_reference_api(codec, cubes, oracle_latents):
    a = model_voxception.AnalysisTransform().bind(codec)
    y = a(cubes[0][:1].astype(np.float64))              # points2voxels hands over float64
    assert y.shape == (1, 16, 16, 16, 16) and y.numpy().dtype == np.float32
    z = model_voxception.HyperEncoder().bind(codec)(y)
    loc, scale = model_voxception.HyperDecoder().bind(codec)(np.rint(z.numpy()))
    assert z.shape == (1, 8, 8, 8, 8) and loc.shape == scale.shape == (1, 16, 16, 16, 16)
    assert (scale.numpy() >= 0).all()
    x = model_voxception.SynthesisTransform().bind(codec)(np.rint(y.numpy()))
    assert x.shape == (1, 64, 64, 64, 1)


def test_determinism_and_batch_invariance(codec, cubes):
    """Bit-reproducible across runs and independent of batch composition (README.md:111-114 is the
    failure this removes)."""
    x = codec.to_device(cubes[0])
    y1 = codec.analysis(x)
    y2 = codec.analysis(x)
    assert torch.equal(y1, y2)
    y_single = torch.cat([codec.analysis(x[i:i + 1]) for i in range(x.shape[0])])
    assert torch.equal(y1, y_single)
    z = torch.round(codec.hyper_encode(y1))
    l1, s1 = codec.hyper_decode(z)
    l2, s2 = codec.hyper_decode(torch.flip(z, [0]))
    assert torch.equal(l1, torch.flip(l2, [0])) and torch.equal(s1, torch.flip(s2, [0]))
    g1 = codec.synthesis(torch.round(y1))
    g2 = codec.synthesis(torch.round(y1)[1:2])
    assert torch.equal(g1[1:2], g2)


def test_simple_model_parity(codec_simple, cubes):
    w = W.synthetic_weights("simple")
    x = cubes[0][:2]
    y = codec_simple.analysis(codec_simple.to_device(x)).cpu().numpy()
    ref = nets.run_net("simple", "analysis", x.astype(np.float32), W.net_weights(w, "analysis_transform"))
    assert y.shape == (2, 8, 8, 8, 32)
    assert np.abs(y - ref).max() < 3e-4 * max(1.0, np.abs(ref).max())
    assert _round_agreement(y, ref)[0] >= 0.999
    logits = codec_simple.synthesis(codec_simple.to_device(np.rint(ref))).cpu().numpy()
    ref_l = nets.run_net("simple", "synthesis", np.rint(ref), W.net_weights(w, "synthesis_transform"))
    assert np.abs(logits - ref_l).max() < 3e-4 * max(1.0, np.abs(ref_l).max())


# ------------------------------------------------------------------------------- entropy models
def test_entropy_bottleneck_golden(codec, golden):
    g = golden("golden_entropy.npz")
    for C in (8, 16, 32):
        p = {k[len("eb%d_" % C):]: v for k, v in g.items() if k.startswith("eb%d_" % C) and k.split("_")[1] in ("matrix", "bais", "factor")}
        codec.load_bottleneck(1, p)
        eb = entropy_model.EntropyBottleneck().bind(codec, 1)
        x = g["eb%d_x" % C]
        x_hat, lik = eb(x, False)
        assert np.array_equal(x_hat.numpy(), g["eb%d_x_hat" % C])
        np.testing.assert_allclose(lik.numpy(), g["eb%d_lik" % C], rtol=LIK_RTOL, atol=LIK_ATOL)
        bits = eb.estimate_bits(x)
        assert abs(bits - entropy.estimated_bits(g["eb%d_lik" % C])) <= 1e-3 * bits
        s, mn, mx = eb.compress(x)
        assert (int(mn), int(mx)) == (int(g["eb%d_min" % C]), int(g["eb%d_max" % C]))
        cdf = eb._get_cdf(int(mn), int(mx))[0]
        ref_cdf = g["eb%d_cdf" % C]
        assert cdf.shape == ref_cdf.shape and np.abs(cdf - ref_cdf).max() <= 2
        assert (cdf[:, 0] == 0).all() and (cdf[:, -1] == 65536).all() and (np.diff(cdf, axis=-1) >= 1).all()
        dec = eb.decompress(s.numpy(), mn.numpy(), mx.numpy(), np.array(x.shape), C)
        assert np.array_equal(dec.numpy(), g["eb%d_x_hat" % C])
    # restore the codec's own 16-channel slot
    w = codec.weights
    codec.load_bottleneck(1, {k[len("estimator_y/"):]: v for k, v in w.items() if k.startswith("estimator_y/")})


def test_symmetric_conditional_golden(codec, golden):
    g = golden("golden_entropy.npz")
    sc = conditional_entropy_model.SymmetricConditional().bind(codec)
    y, loc, scale = g["sc_y"], g["sc_loc"], g["sc_scale"]
    y_hat, lik = sc(y, loc, scale, False)
    assert np.array_equal(y_hat.numpy(), g["sc_y_hat"])
    np.testing.assert_allclose(lik.numpy(), g["sc_lik"], rtol=LIK_RTOL, atol=LIK_ATOL)
    s, mn, mx = sc.compress(y, loc, scale)
    assert (int(mn), int(mx)) == (int(g["sc_min"]), int(g["sc_max"]))
    dec = sc.decompress(s.numpy(), loc, scale, mn.numpy(), mx.numpy(), np.array(y.shape))
    assert np.array_equal(dec.numpy(), g["sc_y_hat"])
    # coded size within a few bytes of the oracle's (tables may differ in the last unit)
    assert abs(len(s.numpy()) - len(g["sc_string"])) <= 8


def test_laplace_likelihood_bits_minmax_vs_oracle(codec, oracle_latents):
    o = oracle_latents
    sc_or = entropy.SymmetricConditionalOracle()
    y, loc, scale = (codec.to_device(o[k]) for k in ("y", "loc", "scale"))
    B = y.shape[0]
    y_hat, p, bits, mm = codec.laplace(y.reshape(B, -1), loc.reshape(B, -1), scale.reshape(B, -1))
    ref_hat, ref_p = sc_or(o["y"], o["loc"], o["scale"])
    assert np.array_equal(y_hat.cpu().numpy().reshape(ref_hat.shape), ref_hat)
    np.testing.assert_allclose(p.cpu().numpy().reshape(ref_p.shape), ref_p, rtol=LIK_RTOL, atol=LIK_ATOL)
    for b in range(B):
        rb = entropy.estimated_bits(ref_p[b])
        assert abs(float(bits[b]) - rb) <= 1e-3 * rb
        assert int(mm[b, 0]) == int(ref_hat[b].min()) and int(mm[b, 1]) == int(ref_hat[b].max())


def test_factorized_likelihood_vs_oracle(codec, oracle_latents):
    o = oracle_latents
    p = {k[len("estimator/"):]: v for k, v in o["w"].items() if k.startswith("estimator/")}
    eb_or = entropy.EntropyBottleneckOracle(p)
    ref_hat, ref_p = eb_or(o["z"])
    z_hat, pz, bits, mm = codec.factorized(0, codec.to_device(o["z"]))
    assert np.array_equal(z_hat.cpu().numpy(), ref_hat)
    np.testing.assert_allclose(pz.cpu().numpy(), ref_p, rtol=LIK_RTOL, atol=LIK_ATOL)
    rb = entropy.estimated_bits(ref_p)
    assert abs(float(bits[0]) - rb) <= 1e-3 * rb
    assert (int(mm[0]), int(mm[1])) == (int(ref_hat.min()), int(ref_hat.max()))


def test_conditional_cdf_rows_vs_oracle(codec, oracle_latents):
    """Per-element quantised CDF rows.  The pmf differs from NumPy's by ulps of expf, which can move a
    16-bit table entry by one unit; rows must be valid, within 2 units, and mostly identical."""
    o = oracle_latents
    sc_or = entropy.SymmetricConditionalOracle()
    b = 0
    y_hat = np.rint(o["y"][b]).reshape(-1)
    mn, mx = int(y_hat.min()), int(y_hat.max())
    N = mx - mn + 1
    loc, scale = o["loc"][b].reshape(-1), o["scale"][b].reshape(-1)
    ref = sc_or.get_cdf(loc, scale, mn, mx)                        # [E, N+1]
    rows, off = codec.laplace_cdf(codec.to_device(loc[None]), codec.to_device(scale[None]), np.array([[mn, mx]], np.int32))
    got = rows.cpu().numpy().view(np.uint16).reshape(-1, N).astype(np.int64)
    assert off[-1] == got.size
    assert (got[:, 0] == 0).all() and (np.diff(got, axis=1) >= 1).all() and (got[:, -1] < 65536).all()
    diff = np.abs(got - ref[:, :N])
    same = (diff.max(axis=1) == 0).mean()
    print("conditional CDF rows identical to the oracle: %.4f, max unit diff %d" % (same, diff.max()))
    assert diff.max() <= 2 and same >= 0.95
    # encoder-side intervals are the same rows looked up at the symbol
    mm = codec.to_device(np.array([[mn, mx]], np.int32))
    iv = codec.laplace_intervals(codec.to_device(y_hat[None]), codec.to_device(loc[None]), codec.to_device(scale[None]), mm)
    iv = iv.cpu().numpy().view(np.uint32).reshape(-1)
    sym = (y_hat - mn).astype(np.int64)
    full = np.concatenate([got, np.full((got.shape[0], 1), 65536)], axis=1)
    lower = full[np.arange(sym.size), sym]
    upper = full[np.arange(sym.size), sym + 1]
    assert np.array_equal(iv & 0xFFFF, lower) and np.array_equal((iv >> 16) + 1, upper - lower)


# ------------------------------------------------------------------------------- codec round trips
def test_hyper_round_trip_and_latent_parity(codec, cubes, oracle_latents):
    c, nums = cubes
    o = oracle_latents
    out = transform.compress_hyper(c, model_voxception, "", decompress=True)
    y_strings, y_min_vs, y_max_vs, y_shape, z_strings, z_min_v, z_max_v, z_shape, x_enc = out
    assert list(y_shape.numpy()) == [1, 16, 16, 16, 16] and list(z_shape.numpy()) == [3, 8, 8, 8, 8]
    assert len(y_strings.numpy()) == 3 and all(isinstance(s, bytes) and len(s) > 0 for s in y_strings.numpy())
    # decode from the byte strings alone, through the file-format types (int8/uint8 headers, np arrays)
    xs = transform.decompress_hyper(np.array(list(y_strings.numpy())), y_min_vs.numpy().astype(np.int32),
                                    y_max_vs.numpy().astype(np.int32), y_shape.numpy().astype(np.int16), z_strings.numpy(),
                                    np.int8(z_min_v.numpy()), np.int8(z_max_v.numpy()), z_shape.numpy().astype(np.int16),
                                    model_voxception, "")
    assert np.array_equal(xs.numpy(), x_enc.numpy()), "decoder-side reconstruction differs from encoder-side"
    # quantised latents against the oracle
    y_gpu = codec.analysis(codec.to_device(c))
    assert _round_agreement(y_gpu.cpu().numpy(), o["y"], o["y64"])[0] >= 0.999
    for b in range(3):
        assert int(y_min_vs.numpy()[b]) == int(np.rint(y_gpu[b].cpu().numpy()).min())
    # decoded voxel set == oracle's top-k on the same logits
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec)
    ref_mask = topk.select_voxels(xs.numpy(), nums, 1.0)
    assert np.array_equal(mask, ref_mask)
    # coded size close to the estimated bits
    bits = conditional_entropy_model.SymmetricConditional().bind(codec).estimate_bits(
        y_gpu, *codec.hyper_decode(torch.round(codec.hyper_encode(y_gpu))))
    for b in range(3):
        assert len(y_strings.numpy()[b]) * 8 <= bits[b] * 1.02 + 64


def test_factorized_round_trip_simple_model(codec_simple, cubes):
    c, nums = cubes
    strings, min_v, max_v, shape = transform.compress_factorized(c, model_simple, "")
    assert list(shape.numpy()) == [3, 8, 8, 8, 32]
    xs = transform.decompress_factorized(strings.numpy(), np.int8(min_v.numpy()), np.int8(max_v.numpy()),
                                         shape.numpy().astype(np.int16), model_simple, "")
    y = codec_simple.analysis(codec_simple.to_device(c))
    ref = codec_simple.synthesis(torch.round(y))
    assert np.array_equal(xs.numpy(), ref.cpu().numpy())
    # oracle decodes the same string to the same symbols (same integer CDF rule, tables from the GPU pmf)
    w = W.synthetic_weights("simple")
    p = {k[len("estimator/"):]: v for k, v in w.items() if k.startswith("estimator/")}
    cdf = codec_simple.factorized_cdf(0, int(min_v), int(max_v))
    n = int(np.prod(shape.numpy()))
    sym = coder.range_decode(strings.numpy(), n, cdf, np.tile(np.arange(32, dtype=np.int32), n // 32))
    assert np.array_equal(sym.astype(np.int32) + int(min_v), np.rint(y.cpu().numpy()).reshape(-1).astype(np.int32))


def test_factorized_round_trip_voxception(codec, cubes):
    c, _ = cubes
    strings, min_v, max_v, shape = transform.compress_factorized(c[:2], model_voxception, "")
    xs = transform.decompress_factorized(strings, min_v, max_v, shape, model_voxception, "")
    ref = codec.synthesis(torch.round(codec.analysis(codec.to_device(c[:2]))))
    assert np.array_equal(xs.numpy(), ref.cpu().numpy())


def test_single_symbol_alphabet_raises(codec):
    """entropy_model.py:192-193 TODO: a one-symbol alphabet cannot be converted to a quantised CDF."""
    eb = entropy_model.EntropyBottleneck().bind(codec, 0)
    with pytest.raises(_lib.PcgcError) as e:
        eb.compress(np.zeros((1, 8, 8, 8, 8), np.float32))
    assert e.value.code == -2
    sc = conditional_entropy_model.SymmetricConditional().bind(codec)
    z = np.zeros((1, 4, 4, 4, 16), np.float32)
    old = conditional_entropy_model.WIDEN_RANGES
    try:
        conditional_entropy_model.WIDEN_RANGES = False                    # the reference's exact (min, max): TF raises here too
        with pytest.raises(_lib.PcgcError):
            sc.compress(z, z, z + 1.0)
    finally:
        conditional_entropy_model.WIDEN_RANGES = old


def test_degenerate_cube_ranges_are_widened_and_round_trip(codec):
    """An all-zero latent cube (plausible for sparse cubes at low rate) and a cube whose latents are all positive: the coded
    range is widened to contain 0 and two symbols, so the cloud does not abort after all GPU work and the header byte of the
    container (max*16 - min, min <= 0 <= max) can hold it; the decoder rebuilds the tables from the header alone."""
    sc = conditional_entropy_model.SymmetricConditional().bind(codec)
    rng = np.random.default_rng(0)
    y = np.zeros((3, 4096), np.float32)
    y[1] = rng.integers(2, 6, 4096)                                       # all positive: range [2, 5] -> [0, 5]
    y[2] = rng.integers(-3, 4, 4096)
    loc = np.zeros_like(y)
    scale = np.ones_like(y)
    yd, ld, sd = codec.to_device(y), codec.to_device(loc), codec.to_device(scale)
    strings, mins, maxs = sc.compress_cubes(yd, ld, sd)
    assert list(mins) == [0, 0, -3] and list(maxs) == [1, 5, 3]
    back = sc.decompress_cubes(strings, ld, sd, mins, maxs)
    assert torch.equal(back, yd)


# ------------------------------------------------------------------------------- top-k
def test_topk_golden(codec, golden):
    g = golden("golden_topk.npz")
    vols, nums = g["vols"], g["nums"]
    mask = inout_points.select_voxels(vols, nums, 1.0, codec=codec)
    assert mask.dtype == np.float32 and mask.shape == vols.shape
    assert np.array_equal(np.packbits(mask.astype(np.uint8)), g["mask"])
    m2 = inout_points.select_voxels(vols[:5], nums[:5], 0.37, codec=codec)
    assert np.array_equal(np.packbits(m2.astype(np.uint8)), g["mask_rho"])
    m3 = inout_points.select_voxels(vols, nums, 1.0, fixed_thres=-1.0, codec=codec)
    assert np.array_equal(np.packbits(m3.astype(np.uint8)), g["mask_fixed"])
    assert np.array_equal(inout_points.voxels2points(mask)[0], g["points0"])


def test_topk_full_size_vs_oracle(codec):
    rng = np.random.default_rng(4)
    vols = rng.normal(-3, 4, (5, 64, 64, 64, 1)).astype(np.float32)
    vols[1] = np.round(vols[1])                       # massive ties
    vols[2] = np.float32(0.25)                        # constant cube
    nums = np.array([4246, 11450, 78, 262144, 1], np.int64)
    mask = inout_points.select_voxels(vols, nums, 1.0, codec=codec)
    assert np.array_equal(mask, topk.select_voxels(vols, nums, 1.0))
    m, thres, cnt = codec.topk(codec.to_device(vols), codec.to_device(nums.astype(np.int32)))
    assert np.array_equal(cnt.cpu().numpy(), mask.reshape(5, -1).sum(1).astype(np.int32))
    for b in range(5):
        assert float(thres[b]) == topk.get_adaptive_thres(vols[b], int(nums[b]))
    with pytest.raises(IndexError):
        inout_points.select_voxels(vols[:1], [262145], 1.0, codec=codec)


# ------------------------------------------------------------------------------- full-size properties
def test_vox10_properties_full_size():
    """BASELINE config 1 at full size (~190 cubes): size-independent properties -- the stream decodes
    to exactly the encoder-side reconstruction, every cube is batch-invariant, symbol ranges fit the
    bitstream format's [-15, 15] / int8 headers (inout_bitstream.py:95-96,111)."""
    codec = runtime.get_codec("voxception", "")
    c, pos, nums = synthetic.workload("vox10")
    out = transform.compress_hyper(c, model_voxception, "", decompress=True)
    y_strings, y_min_vs, y_max_vs, y_shape, z_strings, z_min_v, z_max_v, z_shape, x_enc = out
    assert y_min_vs.numpy().min() >= -15 and y_max_vs.numpy().max() <= 15
    assert -128 <= int(z_min_v) and int(z_max_v) <= 127
    xs = transform.decompress_hyper(y_strings.numpy(), y_min_vs.numpy(), y_max_vs.numpy(), y_shape.numpy(), z_strings.numpy(),
                                    z_min_v.numpy(), z_max_v.numpy(), z_shape.numpy(), model_voxception, "")
    assert torch.equal(xs.tensor, x_enc.tensor)
    i = len(c) // 2
    one = transform.compress_hyper(c[i:i + 1], model_voxception, "")
    assert one[0].numpy()[0] == y_strings.numpy()[i]
    mask = inout_points.select_voxels(xs, nums, 1.0, codec=codec)
    assert (mask.reshape(len(c), -1).sum(1) >= nums).all()


def test_sharded_codec_world1_equals_transform(codec, cubes):
    """pcgcv1_b200.sharding (the multi-GPU form) with a one-rank group reproduces transform.py's
    stream byte for byte and the same voxel masks."""
    import os, socket
    import torch.distributed as dist
    from pcgcv1_b200 import sharding
    c, nums = cubes
    created = False
    if not dist.is_initialized():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)
        created = True
    try:
        local = sharding.GpuLocalCodec("voxception", "")
        stream = sharding.compress_sharded(c, local)
        masks = sharding.decompress_sharded(stream, nums, 1.0, local)
    finally:
        if created:
            dist.destroy_process_group()
    out = transform.compress_hyper(c, model_voxception, "")
    assert stream["y_strings"] == list(out[0].numpy())
    assert stream["z_string"] == out[4].numpy() and (stream["z_min"], stream["z_max"]) == (int(out[5]), int(out[6]))
    xs = transform.decompress_hyper(*[o.numpy() for o in out], model_voxception, "")
    ref = inout_points.select_voxels(xs, nums, 1.0, codec=codec, dtype="uint8")
    assert np.array_equal(masks, ref)


@pytest.mark.parametrize("N", [2, 7, 8, 9, 16, 23, 24, 31, 33, 48, 64])
def test_device_normaliser_equals_host_bit_for_bit(codec, N):
    """The sub-warp (8 lanes per row) device normaliser against the host function of the same library (which the CPU
    tests tie to the oracle's greedy definition): identical on every row, incl. surplus rows, large deficits
    (water-filling) and exact ties."""
    rng = np.random.default_rng(100 + N)
    rows = 4096
    pmf = rng.random((rows, N)).astype(np.float32) ** rng.integers(1, 9, (rows, 1))
    pmf /= pmf.sum(-1, keepdims=True)
    pmf *= rng.choice([1.0, 1.0, 0.999, 0.98, 0.9, 0.5, 0.1, 1.01, 1.2], (rows, 1)).astype(np.float32)
    pmf = np.maximum(pmf, 1e-9).astype(np.float32)
    pmf[:8] = 1.0 / N                                   # exact ties
    pmf[8:16, 1:] = 1e-9; pmf[8:16, 0] = 1.0            # one dominant symbol, the rest at the likelihood bound
    sc_or = entropy.SymmetricConditionalOracle()
    mn = -(N // 2)
    lap = sc_or.pmf(rng.normal(0, 2, 1024).astype(np.float32), (np.abs(rng.normal(0, 1.5, 1024)) + 1e-3).astype(np.float32), mn, mn + N - 1)
    pmf[16:16 + 1024] = lap
    pmf = np.ascontiguousarray(pmf)
    L = codec.lib
    ref = np.empty((rows, N + 1), np.int32)
    assert L.pcgc_pmf_to_quantized_cdf(pmf.ctypes.data, rows, N, 16, ref.ctypes.data) == 0
    d_pmf = codec.to_device(pmf)
    d_cdf = torch.empty((rows, N + 1), dtype=torch.int32, device=codec.dev)
    codec._stream()
    codec._check(L.pcgc_debug_quantize_pmf(codec.ctx, d_pmf.data_ptr(), rows, N, 16, d_cdf.data_ptr()))
    got = d_cdf.cpu().numpy()
    bad = (got != ref).any(-1)
    assert not bad.any(), "first mismatching row %d: %s vs %s" % (np.argmax(bad), np.diff(got[np.argmax(bad)]), np.diff(ref[np.argmax(bad)]))
