"""I/O rows of SURVEY.md section 8(f) (GPU part): voxelisation and ordered point extraction kernels against the NumPy
restatement of points2voxels / voxels2points (inout_points.py:116-143), and the CLI (test.py) round trip through the
on-disk container.  Integer / index work: bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _small_cloud(n_cubes_hint=24):
    from pcgcv1_b200 import synthetic
    pts = synthetic.cloud_vox10()
    sel = pts[(pts[:, 2] > 690) & (pts[:, 0] < 540)]                      # a cap of the surface: a few dozen cubes
    return np.concatenate([sel, sel[:100]])                               # with duplicates


def test_voxelize_matches_points2voxels(codec, tmp_path):
    from pcgcv1_b200.dataprocess import inout_points
    pts = _small_cloud()
    ply = str(tmp_path / "c.ply")
    inout_points.write_ply_data(ply, pts)
    sp, cp = inout_points.load_points(ply, 64, 64)
    local, offsets, cp2 = inout_points.load_points_packed(ply, 64, 64)
    assert np.array_equal(cp, cp2) and len(sp) == len(offsets) - 1 and len(sp) > 4
    want = inout_points.points2voxels(sp, 64)                              # NumPy restatement (uint8, values as the reference)
    got = inout_points.points2voxels_device(local, offsets, 64, codec=codec)
    assert got.shape == want.shape
    assert np.array_equal(got.numpy(), want)
    assert np.array_equal(codec.count_voxels(got.tensor), want.sum(axis=(1, 2, 3, 4)))


def test_voxelize_rejects_out_of_range(codec):
    from pcgcv1_b200._lib import PcgcError
    local = np.array([[1, 2, 3], [64, 0, 0]], np.int16)
    with pytest.raises(PcgcError):
        codec.voxelize(local, np.array([0, 2], np.int64), 64)
    assert codec.voxelize(np.zeros((0, 3), np.int16), np.array([0, 0, 0], np.int64), 64).sum().item() == 0


def test_extract_points_matches_np_where(codec):
    import torch
    from pcgcv1_b200.dataprocess import inout_points
    rng = np.random.default_rng(3)
    B = 7
    mask = (rng.random((B, 64, 64, 64, 1)) < 0.02).astype(np.uint8)
    mask[2] = 0                                                           # an empty cube
    mask[5, :, :, :, 0] = (rng.random((64, 64, 64)) < 0.6) * 3            # dense cube, non-0/1 values count as set
    want = inout_points.voxels2points(mask)                               # np.where order
    md = torch.from_numpy(mask).to(codec.dev)
    for cap in (None, 10, sum(len(w) for w in want)):                     # unknown, too small (retry) and exact capacity
        pts, counts = inout_points.voxels2points_device(md, codec=codec, cap=cap)
        assert counts.tolist() == [len(w) for w in want]
        assert pts.dtype == np.int16 and np.array_equal(pts, np.concatenate(want).astype(np.int16))
    e, c = inout_points.voxels2points_device(torch.zeros((2, 64, 64, 64, 1), dtype=torch.uint8, device=codec.dev), codec=codec)
    assert e.shape == (0, 3) and c.tolist() == [0, 0]


@pytest.mark.parametrize("mode,modelname", [("hyper", "models.model_voxception"), ("factorized", "models.model_simple")])
def test_cli_round_trip(tmp_path, monkeypatch, mode, modelname):
    """python -m pcgcv1_b200.test compress X.ply ; decompress compressed/X  (test.py:74-115) on a small cloud: the files
    exist with the container's layout and the reconstruction equals what the in-memory API gives for the same cubes."""
    import importlib
    from pcgcv1_b200 import runtime, test as cli, transform
    from pcgcv1_b200.dataprocess import inout_bitstream, inout_points
    from pcgcv1_b200.process import preprocess
    monkeypatch.chdir(tmp_path)
    pts = _small_cloud()
    inout_points.write_ply_data("cloud_vox10.ply", pts)
    cli.main(["compress", "cloud_vox10.ply", "--mode", mode, "--modelname", modelname])
    exts = [".strings", ".pointnums", ".cubepos"] + ([".strings_head", ".strings_hyper"] if mode == "hyper" else [])
    for ext in exts:
        assert os.path.getsize(os.path.join("compressed", "cloud_vox10" + ext)) > 0
    cli.main(["decompress", "compressed/cloud_vox10", "--mode", mode, "--modelname", modelname])
    rec = inout_points.load_ply_data("cloud_vox10_rec.ply")

    # the same through the in-memory API
    model = importlib.import_module("pcgcv1_b200." + modelname)
    codec = runtime.get_codec(model, "")
    cubes, cube_positions, points_numbers = preprocess("cloud_vox10.ply", 1.0, 64, 64, codec=codec)
    if mode == "hyper":
        out = transform.compress_hyper(cubes, model, "")
        xs = transform.decompress_hyper(*[o.numpy() for o in out], model, "")
        r = inout_bitstream.read_binary_files_hyper("cloud_vox10", "compressed")
        assert [bytes(s) for s in r[0]] == [bytes(s) for s in out[0].numpy()] and r[1] == out[4].numpy()
    else:
        out = transform.compress_factorized(cubes, model, "")
        xs = transform.decompress_factorized(*[o.numpy() for o in out], model, "")
    mask = inout_points.select_voxels(xs, points_numbers, 1.0, codec=codec)
    want_pts = inout_points.voxels2points(mask)
    inout_points.save_points(want_pts, cube_positions, "want.ply", 64)
    want = inout_points.load_ply_data("want.ply")
    assert len(rec) >= int(points_numbers.sum()) and np.array_equal(rec, want)
    uniq = np.unique(pts, axis=0)
    assert int(points_numbers.sum()) <= len(uniq)


def test_cli_round_trip_with_scaling(tmp_path, monkeypatch):
    """--scale 0.5 (process.py:24-31,72-78): the cloud is down-scaled and de-duplicated before partitioning, the reconstruction is
    scaled back up and written with float coordinates; the temporary scaling files are removed."""
    from pcgcv1_b200 import test as cli
    from pcgcv1_b200.dataprocess import inout_points
    monkeypatch.chdir(tmp_path)
    pts = _small_cloud()
    inout_points.write_ply_data("cloud_vox10.ply", pts)
    cli.main(["compress", "cloud_vox10.ply", "--scale", "0.5", "--min_num", "20"])
    cli.main(["decompress", "compressed/cloud_vox10", "--scale", "0.5"])
    rec = inout_points.load_ply_data("cloud_vox10_rec.ply")
    down = np.unique(np.round(pts.astype("float32") * 0.5), axis=0)
    assert 0.5 * len(down) < len(rec) < 1.5 * len(down)
    assert (rec % 2 == 0).all()                                           # integer grid of the half-scale cloud, scaled by 2
    lo, hi = pts.min(0) - 130, pts.max(0) + 130                           # random-init weights: anywhere inside the kept 64^3 cubes (x2)
    assert (rec >= lo).all() and (rec <= hi).all()
    assert not [f for f in os.listdir(".") if "downscaling" in f or "downsampling" in f]


def test_tf_checkpoint_directory_runs_like_weights_npz(tmp_path, codec):
    """A directory holding a TensorBundle checkpoint (written here in the reference's object-graph naming) loads through
    weights.load -> tf_checkpoint and gives bit-identical transforms and strings to the same weights passed directly."""
    import torch
    from pcgcv1_b200 import runtime, synthetic, tf_checkpoint, transform, weights
    from pcgcv1_b200.models import model_voxception
    w = {k: v for k, v in weights.synthetic_weights("voxception").items() if not k.startswith("estimator_y/")}
    tf_checkpoint.export_checkpoint(str(tmp_path), w, step=3)
    assert not os.path.exists(os.path.join(str(tmp_path), "weights.npz"))
    c, _ = synthetic.surface_cubes(2, seed=4)
    a = transform.compress_hyper(c, model_voxception, "")
    b = transform.compress_hyper(c, model_voxception, str(tmp_path))
    assert [bytes(s) for s in a[0].numpy()] == [bytes(s) for s in b[0].numpy()] and a[4].numpy() == b[4].numpy()
    xa = transform.decompress_hyper(*[o.numpy() for o in a], model_voxception, "")
    xb = transform.decompress_hyper(*[o.numpy() for o in b], model_voxception, str(tmp_path))
    assert torch.equal(xa.tensor, xb.tensor)
