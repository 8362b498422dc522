"""The oracle against the golden vectors produced by running the reference's own source
(tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import torch

from oracle import entropy, nets, topk
from pcgcv1_b200 import weights as W


def _digest(w):
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.digest()


def test_entropy_bottleneck_matches_reference(golden):
    g = golden("golden_entropy.npz")
    for C in (8, 16, 32):
        p = {k[len("eb%d_" % C):]: v for k, v in g.items() if k.startswith("eb%d_" % C) and ("matrix" in k or "bais" in k or "factor" in k)}
        eb = entropy.EntropyBottleneckOracle(p)
        x = g["eb%d_x" % C]
        x_hat, lik = eb(x)
        assert np.array_equal(x_hat, g["eb%d_x_hat" % C])
        np.testing.assert_allclose(lik, g["eb%d_lik" % C], rtol=2e-6, atol=1e-12)
        mn, mx = int(g["eb%d_min" % C]), int(g["eb%d_max" % C])
        np.testing.assert_allclose(eb.pmf(mn, mx), g["eb%d_pmf" % C], rtol=2e-6, atol=1e-12)
        s, mn2, mx2 = eb.compress(x)
        assert (mn2, mx2) == (mn, mx)
        assert np.array_equal(eb.get_cdf(mn, mx), g["eb%d_cdf" % C])
        assert s == g["eb%d_string" % C].tobytes()
        assert np.array_equal(eb.decompress(s, mn, mx, x.shape), x_hat)


def test_symmetric_conditional_matches_reference(golden):
    g = golden("golden_entropy.npz")
    sc = entropy.SymmetricConditionalOracle()
    y, loc, scale = g["sc_y"], g["sc_loc"], g["sc_scale"]
    y_hat, lik = sc(y, loc, scale)
    assert np.array_equal(y_hat, g["sc_y_hat"])
    np.testing.assert_array_equal(lik, g["sc_lik"])          # same NumPy ops in the same order: bit-equal
    mn, mx = int(g["sc_min"]), int(g["sc_max"])
    np.testing.assert_array_equal(sc.pmf(loc.reshape(-1), scale.reshape(-1), mn, mx), g["sc_pmf"])
    s, mn2, mx2 = sc.compress(y, loc, scale)
    assert (mn2, mx2) == (mn, mx) and s == g["sc_string"].tobytes()
    assert np.array_equal(sc.decompress(s, loc, scale, mn, mx, y.shape), y_hat)


def test_transform_graphs_match_reference(golden):
    g = golden("golden_nets.npz")
    wv = W.synthetic_weights("voxception")
    assert _digest(wv) == g["vox_weights_sha256"].tobytes(), "synthetic weights drifted: regenerate the golden vectors"
    cube = g["cube16"]
    y = nets.run_net("voxception", "analysis", cube, W.net_weights(wv, "analysis_transform"))
    np.testing.assert_allclose(y, g["vox_y"], rtol=0, atol=1e-5)
    z = nets.run_net("voxception", "hyper_encoder", g["vox_y"], W.net_weights(wv, "hyper_encoder"))
    np.testing.assert_allclose(z, g["vox_z"], rtol=0, atol=1e-5)
    loc, scale = nets.run_net("voxception", "hyper_decoder", np.rint(g["vox_z"]), W.net_weights(wv, "hyper_decoder"))
    np.testing.assert_allclose(loc, g["vox_loc"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(scale, g["vox_scale"], rtol=0, atol=1e-5)
    x = nets.run_net("voxception", "synthesis", np.rint(g["vox_y"]), W.net_weights(wv, "synthesis_transform"))
    np.testing.assert_allclose(x, g["vox_logits"], rtol=0, atol=1e-5)
    ws = W.synthetic_weights("simple")
    assert _digest(ws) == g["simple_weights_sha256"].tobytes()
    ys = nets.run_net("simple", "analysis", cube, W.net_weights(ws, "analysis_transform"))
    np.testing.assert_allclose(ys, g["simple_y"], rtol=0, atol=1e-4)
    xs = nets.run_net("simple", "synthesis", np.rint(g["simple_y"]), W.net_weights(ws, "synthesis_transform"))
    np.testing.assert_allclose(xs, g["simple_logits"], rtol=0, atol=1e-4)


def test_transposed_conv_is_adjoint_of_same_conv():
    """The SAME / Conv3DTranspose rules restated in oracle/nets.py (they live in TF, not in the
    reference tree): <conv_s2(u), v> == <u, conv_transpose_s2(v)> with the same kernel."""
    rng = np.random.default_rng(0)
    for k in (3, 5, 9):
        n, ci, co = 8, 3, 2
        kern = rng.normal(size=(k, k, k, ci, co))
        u = torch.from_numpy(rng.normal(size=(1, n, n, n, ci)))
        v = torch.from_numpy(rng.normal(size=(1, n // 2, n // 2, n // 2, co)))
        fu = nets.conv3d_same(u, {"c/kernel": kern}, "c", stride=2)
        # Conv3DTranspose kernel layout [k,k,k,Cout,Cin] with Cout = ci here
        ftv = nets.conv3d_transpose_same(v, {"c/kernel": kern}, "c", stride=2)
        assert abs(float((fu * v).sum() - (u * ftv).sum())) < 1e-9


def test_topk_matches_reference(golden):
    g = golden("golden_topk.npz")
    vols, nums = g["vols"], g["nums"]
    shape = vols.shape
    mask = topk.select_voxels(vols, nums, 1.0)
    assert np.array_equal(np.packbits(mask.astype(np.uint8)), g["mask"])
    m2 = topk.select_voxels(vols[:5], nums[:5], 0.37)
    assert np.array_equal(np.packbits(m2.astype(np.uint8)), g["mask_rho"])
    m3 = topk.select_voxels(vols, nums, 1.0, fixed_thres=-1.0)
    assert np.array_equal(np.packbits(m3.astype(np.uint8)), g["mask_fixed"])
    pts = topk.voxels2points(mask)
    assert np.array_equal(pts[0], g["points0"])
    assert np.array_equal(np.packbits(topk.points2voxels([pts[0]], 16).astype(np.uint8)), g["vox_from_points0"])
    # ties are kept: at least k voxels selected
    for i in range(5):
        assert mask[i].sum() >= nums[i]
    assert mask.shape == shape


def test_oracle_convs_match_the_direct_definition():
    """oracle/nets.py (torch conv3d + pad / crop rules) against tests/golden/ref_conv.py, which writes every output straight
    from TensorFlow's documented SAME / conv3d_transpose definitions: all kernel sizes and strides of the two models, odd and
    even extents, float64 so that only the index arithmetic is compared."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_conv
    rng = np.random.default_rng(3)
    for k, s, n in ((3, 1, 6), (3, 1, 5), (1, 1, 4), (3, 2, 8), (3, 2, 6), (5, 2, 8), (9, 2, 8), (5, 2, 4)):
        ci, co = 3, 4
        kern = rng.normal(size=(k, k, k, ci, co))
        bias = rng.normal(size=co)
        x = rng.normal(size=(2, n, n, n, ci))
        a = nets.conv3d_same(torch.from_numpy(x), {"c/kernel": kern, "c/bias": bias}, "c", stride=s).numpy()
        b = ref_conv.conv3d_same(x, kern, bias, stride=s)
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-12, (k, s, n)
    for k, n in ((3, 4), (3, 3), (5, 4), (9, 4), (9, 2)):
        ci, co = 3, 2
        kern = rng.normal(size=(k, k, k, co, ci))
        bias = rng.normal(size=co)
        x = rng.normal(size=(2, n, n, n, ci))
        a = nets.conv3d_transpose_same(torch.from_numpy(x), {"c/kernel": kern, "c/bias": bias}, "c", stride=2).numpy()
        b = ref_conv.conv3d_transpose_same(x, kern, bias, stride=2)
        assert a.shape == b.shape == (2, 2 * n, 2 * n, 2 * n, co) and np.abs(a - b).max() < 1e-12, (k, n)
