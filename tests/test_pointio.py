"""I/O rows of SURVEY.md section 8(f) (CPU part): the PLY reader / writer, the cube partition and the bitstream container
against golden vectors produced by running the reference's own dataprocess/inout_points.py and inout_bitstream.py
(tests/golden/make_golden_io.py).  Bit-exact everywhere: integer / byte work."""
import os

import numpy as np
import pytest

from pcgcv1_b200.dataprocess import inout_bitstream, inout_points
from pcgcv1_b200.myutils import gpcc_wrapper

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_io.npz"), allow_pickle=False)


@pytest.fixture()
def ply(tmp_path):
    p = tmp_path / "cloud.ply"
    p.write_bytes(G["ply_text"].tobytes())
    return str(p)


def test_load_ply_data_matches_reference(ply):
    a = inout_points.load_ply_data(ply)
    assert a.dtype == np.int32 and np.array_equal(a, G["load_ply_data"])


def test_load_ply_data_edge_cases(tmp_path):
    p = tmp_path / "e.ply"
    p.write_text("ply\nend_header\n")
    e = inout_points.load_ply_data(str(p))
    assert e.shape == (0,) and e.dtype == np.int32                      # np.array([]).astype(int32) in the reference
    p.write_text("1 2\n")
    with pytest.raises(IndexError):                                       # wordslist[2] in the reference
        inout_points.load_ply_data(str(p))
    p.write_text("1 2 x\n-1.9 2.9 1e1\r\n+4 5. .5\nnan 1 2\n")
    assert inout_points.load_ply_data(str(p)).tolist() == [[-1, 2, 10], [4, 5, 0], [-2147483648, 1, 2]]


@pytest.mark.parametrize("min_num", [1, 3, 20, 64])
def test_load_points_matches_reference(ply, min_num):
    sp, cp = inout_points.load_points(ply, 64, min_num)
    counts = np.array([1 if p.ndim == 1 else len(p) for p in sp])
    assert np.array_equal(counts, G["lp%d_counts" % min_num])
    assert np.array_equal(np.array([p.ndim for p in sp]), G["lp%d_ndim" % min_num])     # single-point cubes stay 1-D
    assert all(p.dtype == np.int16 for p in sp)
    assert np.array_equal(np.concatenate([p.reshape(-1, 3) for p in sp]), G["lp%d_points" % min_num])
    assert cp.dtype == G["lp%d_cube_positions" % min_num].dtype and np.array_equal(cp, G["lp%d_cube_positions" % min_num])


def test_load_points_nothing_kept_raises(ply):
    with pytest.raises(ValueError):
        inout_points.load_points(ply, 64, 100000)


def test_save_points_and_writer_bytes(ply, tmp_path):
    sp, cp = inout_points.load_points(ply, 64, 20)
    out = tmp_path / "rec.ply"
    inout_points.save_points(sp, cp, str(out), 64)
    assert out.read_bytes() == G["save_points_bytes"].tobytes()
    local, offsets, cp2 = inout_points.load_points_packed(ply, 64, 20)
    inout_points.save_points_packed(local, np.diff(offsets), cp2, str(out), 64)
    assert out.read_bytes() == G["save_points_bytes"].tobytes()
    fl = tmp_path / "f.ply"
    inout_points.write_ply_data(str(fl), G["load_ply_data"][:64].astype("float32") * float(1 / 3))
    assert fl.read_bytes() == G["write_float_bytes"].tobytes()
    # round trip at a larger size with negative coordinates
    rng = np.random.default_rng(1)
    pts = rng.integers(-5000, 5000, size=(200000, 3)).astype(np.int32)
    inout_points.write_ply_data(str(fl), pts)
    assert np.array_equal(inout_points.load_ply_data(str(fl)), pts)


def test_points2voxels_counts_match_reference(ply):
    sp, _ = inout_points.load_points(ply, 64, 20)
    vox = inout_points.points2voxels(sp, 64)
    assert np.array_equal(vox.sum(axis=(1, 2, 3, 4)).astype(np.uint16), G["points_numbers"])


def _hyper_inputs():
    lens = G["hyper_in_y_lens"]
    cat = G["hyper_in_y_concat"].tobytes()
    ends = np.cumsum(lens)
    ys = np.empty(len(lens), dtype=object)
    for i in range(len(lens)):
        ys[i] = cat[ends[i] - lens[i]:ends[i]]
    return ys


def test_bitstream_hyper_bytes_and_round_trip(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                                           # no myutils/tmc3 here: raw cube positions
    ys = _hyper_inputs()
    B = len(ys)
    root = str(tmp_path / "compressed")
    sizes = inout_bitstream.write_binary_files_hyper("g", ys, G["hyper_in_z"].tobytes(), G["points_numbers"], G["cube_positions"],
                                                     G["hyper_in_y_min"], G["hyper_in_y_max"], np.array([1, 16, 16, 16, 16]), np.int32(-9),
                                                     np.int32(11), np.array([B, 8, 8, 8, 8]), rootdir=root)
    for ext in (".strings", ".strings_head", ".strings_hyper", ".pointnums", "_cubepos.ply"):
        assert open(os.path.join(root, "g" + ext), "rb").read() == G["hyper" + ext].tobytes(), ext
    assert sizes[0] == len(G["hyper.strings"]) and sizes[4] == len(gpcc_wrapper.RAW_MAGIC) + 3 * B
    r = inout_bitstream.read_binary_files_hyper("g", root)
    assert [bytes(s) for s in r[0]] == [bytes(s) for s in ys]
    assert r[1] == G["hyper_in_z"].tobytes()
    assert np.array_equal(r[2], G["points_numbers"]) and np.array_equal(r[3], G["cube_positions"])
    assert np.array_equal(r[4], G["hyper_in_y_min"]) and np.array_equal(r[5], G["hyper_in_y_max"])
    assert r[6].tolist() == [1, 16, 16, 16, 16] and (int(r[7]), int(r[8])) == (-9, 11) and r[9].tolist() == [B, 8, 8, 8, 8]


def test_bitstream_factorized_bytes_and_round_trip(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    B = len(G["points_numbers"])
    root = str(tmp_path / "compressed")
    inout_bitstream.write_binary_files_factorized("f", G["fact_in_string"].tobytes(), G["points_numbers"], G["cube_positions"], np.int32(-20),
                                                  np.int32(17), np.array([B, 8, 8, 8, 32]), rootdir=root)
    for ext in (".strings", ".pointnums", "_cubepos.ply"):
        assert open(os.path.join(root, "f" + ext), "rb").read() == G["fact" + ext].tobytes(), ext
    s, pn, cp, mn, mx, shape = inout_bitstream.read_binary_files_factorized("f", root)
    assert s == G["fact_in_string"].tobytes() and np.array_equal(pn, G["points_numbers"]) and np.array_equal(cp, G["cube_positions"])
    assert (int(mn), int(mx)) == (-20, 17) and shape.tolist() == [B, 8, 8, 8, 32]


def test_bitstream_limits_are_checked(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    ys = np.empty(1, dtype=object)
    ys[0] = b"x" * 40000                                                  # > int16: the unsigned extension
    args = dict(points_numbers=[5], cube_positions=np.array([[1, 2, 3]]), y_shape=[1, 16, 16, 16, 16], z_min_v=0, z_max_v=1,
                z_shape=[1, 8, 8, 8, 8], rootdir=str(tmp_path / "c"))
    inout_bitstream.write_binary_files_hyper("a", ys, b"z", y_min_vs=[-3], y_max_vs=[4], **args)
    r = inout_bitstream.read_binary_files_hyper("a", str(tmp_path / "c"))
    assert len(r[0][0]) == 40000 and r[4].tolist() == [-3] and r[5].tolist() == [4]
    with pytest.raises(ValueError):
        inout_bitstream.write_binary_files_hyper("a", ys, b"z", y_min_vs=[-16], y_max_vs=[4], **args)
    ys[0] = b"x" * 70000
    with pytest.raises(ValueError):
        inout_bitstream.write_binary_files_hyper("a", ys, b"z", y_min_vs=[-3], y_max_vs=[4], **args)
    with pytest.raises(ValueError):
        inout_bitstream.write_binary_files_factorized("a", b"s", [5], np.array([[300, 0, 0]]), 0, 1, [1, 8, 8, 8, 32], rootdir=str(tmp_path / "c"))
