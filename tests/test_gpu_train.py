"""Training step on the GPU (BASELINE config 5) against the torch-autograd FP64 oracle of
``train_hyper.py:184-214`` (oracle/train.py).  Tolerances: loss terms 5e-5 relative, every gradient tensor within 2e-3 of its own
scale in the max norm (the CUDA path is exact FP32 with fixed reduction orders; the oracle runs the transforms in float64 and the
likelihood formulas in float32 like the reference -- see oracle/train.py:forward_backward for why that matters)."""
import math

import numpy as np
import pytest
import torch

from oracle import nets, train as otrain
from pcgcv1_b200 import netspec, runtime, synthetic, training, weights as W

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("layer,n", [
    (netspec.Layer("c", 16, 4, 3, 1), 16), (netspec.Layer("c", 16, 4, 1, 1), 16), (netspec.Layer("c", 4, 8, 3, 1, relu=False), 8),
    (netspec.Layer("c", 1, 16, 3, 1), 16), (netspec.Layer("c", 16, 1, 3, 1, relu=False), 16),
    (netspec.Layer("c", 16, 32, 3, 2, bias=False), 16), (netspec.Layer("c", 32, 16, 3, 2, transposed=True), 8),
    (netspec.Layer("c", 64, 16, 3, 1, relu=False), 8),
])
def test_conv_forward_dgrad_wgrad_vs_autograd(codec, layer, n):
    """Every layer family of the model: y, dL/dx, dL/dw, dL/db for a random upstream gradient, vs torch autograd in float64."""
    rng = np.random.default_rng(layer.cin * 100 + layer.cout + layer.k + layer.stride + n)
    B = 2
    x = rng.normal(size=(B, n, n, n, layer.cin)).astype(np.float32)
    kern = (rng.normal(size=netspec.kernel_shape(layer)) / math.sqrt(27 * layer.cin)).astype(np.float32)
    bias = rng.normal(size=layer.cout).astype(np.float32) if layer.bias else None
    ops = training._Ops(codec)
    xd = torch.tensor(x, device=codec.dev, requires_grad=True)
    wd = torch.tensor(kern, device=codec.dev, requires_grad=True)
    bd = torch.tensor(bias, device=codec.dev, requires_grad=True) if layer.bias else None
    y = training._Conv.apply(xd, wd, bd, ops, layer)
    g = rng.normal(size=tuple(y.shape)).astype(np.float32)
    y.backward(torch.tensor(g, device=codec.dev))
    codec.synchronize()
    # oracle
    xo = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wo = {"c/kernel": torch.tensor(kern, dtype=torch.float64, requires_grad=True)}
    if layer.bias:
        wo["c/bias"] = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
    fn = nets.conv3d_transpose_same if layer.transposed else nets.conv3d_same
    yo = fn(xo, wo, "c", stride=layer.stride, relu=layer.relu)
    yo.backward(torch.tensor(g, dtype=torch.float64))
    assert _rel(y.detach().cpu().numpy(), yo.detach().numpy()) < 2e-5
    assert _rel(xd.grad.cpu().numpy(), xo.grad.numpy()) < 2e-5
    assert _rel(wd.grad.cpu().numpy(), wo["c/kernel"].grad.numpy()) < 5e-5
    if layer.bias:
        assert _rel(bd.grad.cpu().numpy(), wo["c/bias"].grad.numpy()) < 5e-5


@pytest.fixture(scope="module")
def step(codec):
    w = W.synthetic_weights("voxception")
    cubes, _ = synthetic.surface_cubes(1, seed=4)
    tr = training.HyperTrainer(codec, w)
    out = tr.forward_backward(cubes, seed=3)
    terms = tr.loss_terms(out)
    grads = {k: v.grad.detach().cpu().numpy().copy() for k, v in tr.params.items() if v.grad is not None}
    ref_terms, ref_grads, ref_act = otrain.forward_backward(w, cubes, seed=3)
    return dict(w=w, cubes=cubes, tr=tr, terms=terms, grads=grads, ref_terms=ref_terms, ref_grads=ref_grads, out=out, ref_act=ref_act)


def test_training_step_loss_vs_oracle(step):
    for k in ("zeros", "ones", "distortion", "bpp_ae", "bpp_hyper", "loss"):
        a, b = step["terms"][k], step["ref_terms"][k]
        print("%-10s gpu %.8g oracle %.8g rel %.2e" % (k, a, b, abs(a - b) / abs(b)))
        assert abs(a - b) <= 5e-5 * abs(b)
    assert _rel(step["out"]["x_tilde"].cpu().numpy(), step["ref_act"]["x_tilde"]) < 1e-3


def test_training_step_gradients_vs_oracle(step):
    g, r = step["grads"], step["ref_grads"]
    worst = ("", 0.0)
    n = 0
    for key, ref in r.items():
        if key.startswith("estimator/"):
            continue
        assert key in g, key
        e = _rel(g[key], ref)
        worst = max(worst, (key, e), key=lambda t: t[1])
        n += 1
        assert e < 2e-3, (key, e)
    print("%d conv gradient tensors, worst relative error %.2e (%s)" % (n, worst[1], worst[0]))
    assert n == 2 * 98 + 3 + 5 + 3 + 5 - 2 - 2 + 0 or n > 200        # every kernel and every bias of the four nets
    # EntropyBottleneck variables: the trainer keeps them concatenated in the checkpoint order
    for name in ("matrix", "bais", "factor"):
        ref = np.concatenate([r["estimator/%s_%d" % (name, i)].reshape(-1) for i in range(4)])
        assert _rel(g["estimator/" + name], ref) < 2e-3, name


def test_training_step_is_deterministic_and_adam_matches(step, codec):
    tr, cubes = step["tr"], step["cubes"]
    tr.forward_backward(cubes, seed=3)
    again = {k: v.grad.detach().cpu().numpy() for k, v in tr.params.items() if v.grad is not None}
    for k, v in again.items():
        assert np.array_equal(v, step["grads"][k]), k                     # bit-identical gradients run to run
    key = "synthesis_transform/deconv_out/kernel"
    p0 = tr.params[key].detach().cpu().numpy().astype(np.float64)
    tr.adam_step()
    tr.adam_step()
    codec.synchronize()
    g = step["grads"][key].astype(np.float64)
    p, m, v = p0, np.zeros_like(p0), np.zeros_like(p0)
    for t in (1, 2):
        p, m, v = otrain.adam_update(p, g, m, v, t, lr=tr.lr)
    got = tr.params[key].detach().cpu().numpy()
    assert np.abs(got - p).max() <= 1e-6 * np.abs(p).max() + 1e-9
    assert np.abs(got - p0).max() > 0
    # a different seed draws different noise
    out2 = tr.forward_backward(cubes, seed=4)
    assert abs(tr.loss_terms(out2)["bpp_ae"] - step["terms"]["bpp_ae"]) > 0


def test_export_weights_round_trip(step):
    tr = step["tr"]
    w2 = tr.export_weights()
    for k in step["w"]:
        if k.split("/")[0] in ("analysis_transform", "synthesis_transform", "hyper_encoder", "hyper_decoder", "estimator"):
            assert k in w2 and w2[k].shape == np.asarray(step["w"][k]).shape, k


def test_focal_loss_kernels_vs_reference_golden_and_autograd(codec):
    """pcgc_train_focal against golden_loss.npz (= /root/reference/loss.py:83-93 executed by tests/golden/make_golden_loss.py, fed
    with float32 sigmoid(pred)), its backward against torch autograd of the oracle's restatement in float64."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_loss.npz"))
    for tag in ("a", "b"):
        pred, label = g[tag + "_pred"], g[tag + "_label"]
        x = torch.tensor(pred, device=codec.dev).contiguous()
        lab = torch.tensor(label.astype(np.uint8), device=codec.dev).contiguous()
        for key, gamma, alpha in (("_focal", 2.0, 0.9), ("_focal_g3_a75", 3.0, 0.75)):
            sums = torch.empty(2, dtype=torch.float64, device=codec.dev)
            grad = torch.empty_like(x)
            codec._stream()
            codec._check(codec.lib.pcgc_train_focal(codec.ctx, x.data_ptr(), lab.data_ptr(), x.numel(), gamma, alpha, sums.data_ptr()))
            codec._check(codec.lib.pcgc_train_focal_backward(codec.ctx, x.data_ptr(), lab.data_ptr(), x.numel(), gamma, alpha, 0.75, grad.data_ptr()))
            codec.synchronize()
            got = float(sums.sum().item())
            assert abs(got - g[tag + key][0]) <= 5e-6 * g[tag + key][0], (tag, key, got, g[tag + key][0])
            xo = torch.tensor(pred, dtype=torch.float64, requires_grad=True)
            f1, f0 = otrain.focal_loss(torch.sigmoid(xo), torch.tensor(label, dtype=torch.float64), gamma, alpha)
            (0.75 * (f1 + f0)).backward()
            ref = xo.grad.numpy()
            # voxels whose probability sits within float32 rounding of a clip bound may fall on either side of it
            p64 = 1.0 / (1.0 + np.exp(-pred.astype(np.float64)))
            clear = (np.abs(p64 - 1e-3) > 1e-6) & (np.abs(p64 - 0.999) > 1e-6)
            assert np.abs(grad.cpu().numpy() - ref)[clear].max() <= 2e-5 * np.abs(ref).max(), (tag, key)
            assert np.count_nonzero(ref == 0) > 0 or tag == "a"        # case b holds clipped voxels: zero gradient there


def test_training_step_with_focal_distortion_vs_oracle(codec):
    w = W.synthetic_weights("voxception")
    cubes, _ = synthetic.surface_cubes(1, seed=4)
    tr = training.HyperTrainer(codec, w, distortion="focal")
    out = tr.forward_backward(cubes, seed=3)
    terms = tr.loss_terms(out)
    grads = {k: v.grad.detach().cpu().numpy().copy() for k, v in tr.params.items() if v.grad is not None}
    ref_terms, ref_grads, _ = otrain.forward_backward(w, cubes, seed=3, distortion="focal")
    for k in ("focal_full", "focal_empty", "distortion", "bpp_ae", "bpp_hyper", "loss"):
        a, b = terms[k], ref_terms[k]
        print("%-12s gpu %.8g oracle %.8g rel %.2e" % (k, a, b, abs(a - b) / abs(b)))
        assert abs(a - b) <= 5e-5 * abs(b)
    worst = ("", 0.0)
    for key, ref in ref_grads.items():
        if key.startswith("estimator/"):
            continue
        e = _rel(grads[key], ref)
        worst = max(worst, (key, e), key=lambda t: t[1])
        assert e < 2e-3, (key, e)
    print("focal: worst relative gradient error %.2e (%s)" % (worst[1], worst[0]))
    with pytest.raises(ValueError):
        training.HyperTrainer(codec, w, distortion="dice")
