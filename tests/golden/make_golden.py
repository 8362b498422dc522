"""Generate the committed golden vectors by EXECUTING THE REFERENCE'S OWN SOURCE in this container.

    python tests/golden/make_golden.py          # needs /root/reference (read-only), writes *.npz here

* ``dataprocess/inout_points.py`` is pure NumPy and is imported as is (select_voxels,
  get_adaptive_thres, voxels2points, points2voxels).
* ``models/entropy_model.py``, ``models/conditional_entropy_model.py``, ``models/model_voxception.py``
  and ``models/model_simple.py`` are imported unmodified with ``tf_shim`` registered as
  ``tensorflow`` (TensorFlow 1.13 cannot be installed here).  See tf_shim.py for what that pins.

/root/reference does not exist on the GPU box: tests only read the .npz files written here.
"""
from __future__ import annotations

import hashlib
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PCGC_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf_shim  # noqa: E402

from pcgcv1_b200 import weights as W  # noqa: E402  (seeded synthetic weights: shared INPUT, not an implementation)


def weights_digest(w) -> str:
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()


def main():
    tf_shim.install()
    sys.path.insert(0, REF)
    ref_points = importlib.import_module("dataprocess.inout_points")
    ref_eb = importlib.import_module("models.entropy_model")
    ref_sc = importlib.import_module("models.conditional_entropy_model")
    ref_vox = importlib.import_module("models.model_voxception")
    ref_simple = importlib.import_module("models.model_simple")
    rng = np.random.default_rng(20191007)

    # ---------------- EntropyBottleneck (models/entropy_model.py) ----------------
    out = {}
    for C in (8, 16, 32):
        p = W.entropy_bottleneck_params(C, np.random.default_rng(100 + C))
        tf_shim.WEIGHTS.clear()
        tf_shim.WEIGHTS.update(p)
        eb = ref_eb.EntropyBottleneck()
        x = (rng.normal(0, 2.5, (2, 3, 4, 5, C))).astype(np.float32)
        x[0, 0, 0, 0, :] = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 3.5, -2.5, 0.0] * (C // 8), np.float32)   # ties: half to even
        x_hat, lik = eb(x, False)
        s, mn, mx = eb.compress(x)
        pmf = tf_shim.CAPTURE["pmf"].copy()
        cdf = eb._get_cdf(mn, mx)
        dec = eb.decompress(s, mn, mx, np.array(x.shape), C)
        assert np.array_equal(dec, x_hat)
        for k, v in p.items():
            out["eb%d_%s" % (C, k)] = v
        out.update({"eb%d_x" % C: x, "eb%d_x_hat" % C: x_hat, "eb%d_lik" % C: lik, "eb%d_min" % C: np.int32(mn),
                    "eb%d_max" % C: np.int32(mx), "eb%d_pmf" % C: pmf.reshape(C, -1), "eb%d_cdf" % C: np.asarray(cdf).reshape(C, -1),
                    "eb%d_string" % C: np.frombuffer(s, np.uint8)})
    # ---------------- SymmetricConditional (models/conditional_entropy_model.py) ----------------
    sc = ref_sc.SymmetricConditional()
    y = rng.normal(0, 2.0, (1, 4, 4, 4, 16)).astype(np.float32)
    loc = rng.normal(0, 1.5, y.shape).astype(np.float32)
    scale = np.maximum(np.abs(rng.normal(0, 0.8, y.shape)).astype(np.float32), 1e-9)
    scale.reshape(-1)[:8] = [1e-9, 1e-4, 1e-2, 0.1, 1.0, 3.0, 10.0, 50.0]      # extreme scales
    loc.reshape(-1)[8:12] = [0.0, 0.5, -0.5, 2.0]                                # sign(2x-loc) corner cases
    y.reshape(-1)[8:12] = [0.0, 0.25, -0.25, 1.0]
    y_hat, lik = sc(y, loc, scale, False)
    s, mn, mx = sc.compress(y, loc, scale)
    pmf = tf_shim.CAPTURE["pmf"].copy()
    dec = sc.decompress(s, loc, scale, mn, mx, np.array(y.shape))
    assert np.array_equal(dec, y_hat)
    out.update({"sc_y": y, "sc_loc": loc, "sc_scale": scale, "sc_y_hat": y_hat, "sc_lik": lik, "sc_min": np.int32(mn),
                "sc_max": np.int32(mx), "sc_pmf": pmf.reshape(-1, pmf.shape[-1]), "sc_string": np.frombuffer(s, np.uint8)})
    np.savez_compressed(os.path.join(HERE, "golden_entropy.npz"), **out)

    # ---------------- transforms (models/model_voxception.py, models/model_simple.py) ----------------
    out = {}
    cube = np.zeros((1, 16, 16, 16, 1), np.float32)
    pts = rng.integers(0, 16, (220, 3))
    cube[0, pts[:, 0], pts[:, 1], pts[:, 2], 0] = 1.0
    out["cube16"] = cube
    wv = W.synthetic_weights("voxception")
    out["vox_weights_sha256"] = np.frombuffer(bytes.fromhex(weights_digest(wv)), np.uint8)

    def use(net):
        tf_shim.WEIGHTS.clear()
        tf_shim.WEIGHTS.update(W.net_weights(wv, net))

    use("analysis_transform");  yv = ref_vox.AnalysisTransform()(cube)
    use("hyper_encoder");       zv = ref_vox.HyperEncoder()(yv)
    use("hyper_decoder");       locv, scalev = ref_vox.HyperDecoder()(np.rint(zv))
    use("synthesis_transform"); xv = ref_vox.SynthesisTransform()(np.rint(yv))
    out.update({"vox_y": yv, "vox_z": zv, "vox_loc": locv, "vox_scale": scalev, "vox_logits": xv})

    ws = W.synthetic_weights("simple")
    out["simple_weights_sha256"] = np.frombuffer(bytes.fromhex(weights_digest(ws)), np.uint8)
    tf_shim.WEIGHTS.clear(); tf_shim.WEIGHTS.update(W.net_weights(ws, "analysis_transform"))
    ys = ref_simple.AnalysisTransform()(cube)
    tf_shim.WEIGHTS.clear(); tf_shim.WEIGHTS.update(W.net_weights(ws, "synthesis_transform"))
    xs = ref_simple.SynthesisTransform()(np.rint(ys))
    out.update({"simple_y": ys, "simple_logits": xs})
    np.savez_compressed(os.path.join(HERE, "golden_nets.npz"), **out)

    # ---------------- top-k (dataprocess/inout_points.py) ----------------
    out = {}
    vols = (rng.random((6, 16, 16, 16, 1)) * 100 - 50).astype(np.float32)
    vols[1] = np.round(vols[1] / 10) * 10          # heavy ties
    vols[2] = -np.abs(vols[2]) - 3.0               # everything below init_thres -> fallback to all voxels
    vols[3].reshape(-1)[:100] = 7.25               # ties straddling the k-th position
    vols[5] = np.round(vols[5])                    # ties incl. +-0
    nums = np.array([1000, 200, 50, 60, 4096, 0], np.int64)
    mask = ref_points.select_voxels(vols, nums, 1.0)
    mask_rho = ref_points.select_voxels(vols[:5], nums[:5], 0.37)
    mask_fixed = ref_points.select_voxels(vols, nums, 1.0, fixed_thres=-1.0)
    pts_list = ref_points.voxels2points(mask)
    out.update({"vols": vols, "nums": nums, "mask": np.packbits(mask.astype(np.uint8)), "mask_rho": np.packbits(mask_rho.astype(np.uint8)),
                "mask_fixed": np.packbits(mask_fixed.astype(np.uint8)), "points0": pts_list[0].astype(np.int16),
                "vox_from_points0": np.packbits(ref_points.points2voxels([pts_list[0]], 16).astype(np.uint8))})
    np.savez_compressed(os.path.join(HERE, "golden_topk.npz"), **out)
    for f in ("golden_entropy.npz", "golden_nets.npz", "golden_topk.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
