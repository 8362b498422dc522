"""The rho search (eval_ablation_studies.py:152-205) and the distortion figures behind it (myutils/pc_error_wrapper.py).

golden_rho.json = the reference's own two functions executed on scripted PSNR sequences (tests/golden/make_golden_rho.py).  The
built-in D1 / D2 figures are checked against a brute-force restatement of their definitions (the MPEG binary is absent: parity of
the figures against the tool itself is unpinned, see the wrapper's docstring)."""
import configparser
import json
import os

import numpy as np
import pytest

from pcgcv1_b200 import eval_ablation_studies as ev
from pcgcv1_b200.myutils import pc_error_wrapper as pw

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_rho.json")))


@pytest.mark.parametrize("case", GOLD["select"], ids=lambda c: "n%d" % len(c["psnr"]))
def test_select_optimal_rho_matches_reference_decisions(case):
    calls, it = [], iter(case["psnr"])
    best = ev.select_optimal_rho("item", case["rhos"], "in.ply", "out.ply", "in_n.ply", None, None, None, 1.0, 64, 1024,
                                 post=lambda f, c, n, p, s, cs, rho: calls.append(rho), metric=lambda a, b, n, res, show=False: {"item": next(it)})
    assert best == case["optimal_rho"]
    assert calls == case["evaluated"]


def test_cfg_post_process_matches_reference(tmp_path):
    case = GOLD["cfg"][0]
    calls, it = [], iter(case["psnr_d1"] + case["psnr_d2"])
    cfg = configparser.ConfigParser()
    cfg["R1"] = {"scale": "1.0"}
    ini = str(tmp_path / "cloud.ini")
    metric = lambda a, b, n, res, show=False: dict.fromkeys((ev.ITEM_D1, ev.ITEM_D2), next(it))
    r1, r2 = ev.cfg_post_process(cfg, ini, "R1", "in.ply", "out.ply", "in_n.ply", None, None, None, 1.0, 64, 1024,
                                 post=lambda f, c, n, p, s, cs, rho: calls.append(rho), metric=metric)
    assert (r1, r2) == (case["rho_d1"], case["rho_d2"])
    assert calls == case["evaluated"]
    assert open(ini).read() == case["ini"]
    # a second call reads the stored values and evaluates nothing
    calls.clear()
    assert ev.cfg_post_process(cfg, ini, "R1", "in.ply", "out.ply", "in_n.ply", None, None, None, 1.0, 64, 1024,
                               post=lambda *a: calls.append(a), metric=None) == (r1, r2)
    assert calls == []


def _brute(a, b, na, peak):
    d = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    ia, ib = d.argmin(1), d.argmin(0)
    e1, e2 = a - b[ia], b - a[ib]
    pp1, pp2 = (e1 ** 2).sum(1), (e2 ** 2).sum(1)
    pl1, pl2 = (e1 * na).sum(1) ** 2, (e2 * na[ib]).sum(1) ** 2
    psnr = lambda v: 10 * np.log10(3 * peak * peak / v)
    return {"mse1      (p2point)": pp1.mean(), "mse2      (p2point)": pp2.mean(), "mseF      (p2point)": max(pp1.mean(), pp2.mean()),
            "mseF,PSNR (p2point)": psnr(max(pp1.mean(), pp2.mean())), "h.        (p2point)": max(pp1.max(), pp2.max()),
            "mseF      (p2plane)": max(pl1.mean(), pl2.mean()), "mseF,PSNR (p2plane)": psnr(max(pl1.mean(), pl2.mean())),
            "h.,PSNR   (p2plane)": psnr(max(pl1.max(), pl2.max()))}


def test_geometry_metrics_against_brute_force_and_files(tmp_path):
    rng = np.random.default_rng(5)
    a = np.unique(rng.integers(0, 64, (400, 3)), axis=0).astype(np.float64)
    b = a[::2] + rng.uniform(-1.2, 1.2, a[::2].shape)          # off-grid: no equidistant neighbours (the plane figures depend on WHICH one)
    na = rng.normal(size=a.shape)
    na /= np.linalg.norm(na, axis=1, keepdims=True)
    got = pw.geometry_metrics(a, b, na, peak=1023.0)
    assert set(got) == set(pw.HEADERS)
    for k, v in _brute(a, b, na, 1023.0).items():
        assert abs(got[k] - v) <= 1e-9 * max(abs(v), 1.0), k
    assert set(pw.geometry_metrics(a, b, None)) == {h for h in pw.HEADERS if "p2point" in h}
    assert pw.geometry_metrics(a, a, na)["mseF,PSNR (p2point)"] == float("inf")
    # through files, as the search calls it: A with normals, B plain
    fa, fb, fn = (str(tmp_path / n) for n in ("a.ply", "b.ply", "a_n.ply"))
    def write(name, pts, normals=None):
        with open(name, "w") as f:
            f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n" % len(pts))
            if normals is not None:
                f.write("property float nx\nproperty float ny\nproperty float nz\n")
            f.write("end_header\n")
            for i, p in enumerate(pts):
                f.write(" ".join(repr(float(x)) for x in (list(p) + (list(normals[i]) if normals is not None else []))) + "\n")
    write(fa, a); write(fb, b); write(fn, a, na)
    assert pw.get_points_number(fa) == len(a)
    if not pw.have_pc_error():
        row = pw.pc_error(fa, fb, fn, 1024)
        for k, v in got.items():
            assert abs(float(row[k].iloc[0]) - v) <= 1e-9 * max(abs(v), 1.0), k


@pytest.mark.gpu
def test_rho_search_end_to_end_on_decoded_cubes(tmp_path):
    """compress -> decompress -> select_optimal_rho over the device-resident logits, built-in D1 figure (binary absent on the box)."""
    from pcgcv1_b200 import synthetic, transform
    from pcgcv1_b200.dataprocess import inout_points as iop
    from pcgcv1_b200.models import model_voxception
    cubes, pos, nums = synthetic.workload("vox10", seed=0, max_cubes=12)
    pos = np.asarray(pos)[:len(cubes)]
    src = str(tmp_path / "in.ply")
    pts = np.concatenate([np.argwhere(cubes[i, ..., 0] > 0) + pos[i] * 64 for i in range(len(cubes))]).astype(np.int32)
    iop.write_ply_data(src, pts)
    out = transform.compress_hyper(cubes, model_voxception, "")
    xs = transform.decompress_hyper(*[o.numpy() for o in out], model_voxception, "")
    rec = str(tmp_path / "rec.ply")
    seen = []
    def metric(a, b, n, res, show=False):
        r = pw.pc_error(a, b, n, res, show)
        seen.append(float(r[ev.ITEM_D1].iloc[0]))
        return r
    rhos = [0.8, 1.0, 1.2, 1.5, 2.0]
    best = ev.select_optimal_rho(ev.ITEM_D1, rhos, src, rec, "", xs, nums[:len(cubes)], pos, 1, 64, 1024, metric=metric)
    # the reference's control flow on the PSNRs actually measured
    mx, want = 0.0, rhos[0]
    for i, p in enumerate(seen):
        mx = 0.0 if i == 0 else max(p, mx)
        if p < mx:
            break
        want = rhos[i]
    assert best == want and best in rhos and len(seen) >= 2 and all(np.isfinite(seen))
    # the file left behind is the last candidate's reconstruction: rho * N points per cube
    n_last = pw.get_points_number(rec)
    assert n_last == sum(int(rhos[len(seen) - 1] * int(n)) for n in nums[:len(cubes)])
