"""The training loop around the GPU step (pcgcv1_b200/train_hyper.py; reference train_hyper.py:100-268): host-side pieces on
CPU, the loop itself (checkpoint, resume, evaluation) on the GPU."""
import json
import os
import random

import numpy as np
import pytest

from pcgcv1_b200 import train_hyper as TH


def _write_cubes(tmp_path, n, seed=0):
    from pcgcv1_b200 import synthetic
    cubes, _ = synthetic.surface_cubes(n, seed=seed)
    files = []
    for i in range(n):
        pts = np.array(np.where(cubes[i, ..., 0] > 0)).T.astype(np.int16)
        f = os.path.join(str(tmp_path), "cube_%03d.npy" % i)
        np.save(f, pts)
        files.append(f)
    return files, cubes


def test_file_list_split_and_sampling_follow_the_reference(tmp_path):
    files = ["f%03d.npy" % i for i in range(90)]
    d = TH.CubeFiles(files)
    assert d.eval_list == files[:10] and d.train_list == files[10:]          # file_list[:len//RATIO_EVAL] is held out (:167,257)
    ref = random.Random(3)                                                    # random.seed(3) (:17), random.sample (:171)
    assert d.sample_train(8) == ref.sample(files[10:], 8)
    assert d.sample_train(8) == ref.sample(files[10:], 8)
    assert TH.CubeFiles(files, rank=1).sample_train(8) != TH.CubeFiles(files, rank=0).sample_train(8)


def test_cube_files_to_voxels_round_trip(tmp_path):
    files, cubes = _write_cubes(tmp_path, 3)
    vox = TH.CubeFiles.voxels(files)
    assert vox.dtype == np.uint8 and np.array_equal(vox, cubes)
    z = os.path.join(str(tmp_path), "c.npz")
    np.savez(z, data=TH.load_cube_points(files[0]))
    assert np.array_equal(TH.load_cube_points(z), TH.load_cube_points(files[0]))


def test_iou_is_get_classify_metrics():
    rng = np.random.default_rng(0)
    pred = (rng.random((2, 8, 8, 8, 1)) > 0.6).astype(np.float32)
    label = (rng.random((2, 8, 8, 8, 1)) > 0.6).astype(np.float32)
    tp = np.sum((pred > 0) & (label > 0)); fp = np.sum((pred > 0) & (label == 0)); fn = np.sum((pred == 0) & (label > 0))
    assert TH.iou(pred, label) == pytest.approx(tp / (tp + fp + fn))              # loss.py:60-78
    assert TH.iou(label, label) == 1.0


@pytest.mark.gpu
def test_training_loop_checkpoints_resumes_and_evaluates(tmp_path):
    import torch
    from pcgcv1_b200 import runtime, weights as W
    files, _ = _write_cubes(tmp_path, 27, seed=11)
    kw = dict(alpha=0.75, beta=3.0, lr=1e-4, batch_size=2, display_step=2, eval_cubes=2, log=lambda *a: None)
    # 4 steps in one go, checkpoint (with the optimizer state) after step 2 and 4
    a_dir = os.path.join(str(tmp_path), "run_a")
    tr_a = TH.train(files, a_dir, num_iteration=4, save_step=2, reset_optimizer=1, **kw)
    wa = tr_a.export_weights()
    assert os.path.exists(os.path.join(a_dir, "weights.npz")) and int(TH.load_train_state(a_dir)["global_step"]) == 4
    log = [json.loads(l) for l in open(os.path.join(a_dir, "log.jsonl"))]
    assert [r["event"] for r in log] == ["train", "eval", "train", "eval"]
    assert all(np.isfinite(r["bpp_ae"]) and 0.0 <= r["IoU"] <= 1.0 for r in log)
    # the same 4 steps as 2 + (resume) 2: the run's own directory is picked up, sampling and noise continue -> identical weights
    b_dir = os.path.join(str(tmp_path), "run_b")
    TH.train(files, b_dir, num_iteration=2, save_step=2, reset_optimizer=1, **kw)
    tr_b = TH.train(files, b_dir, num_iteration=4, save_step=2, reset_optimizer=1, **kw)
    wb = tr_b.export_weights()
    for k in wa:
        assert np.array_equal(wa[k], wb[k]), k
    # the checkpoint is a codec checkpoint: the trained model compresses and decompresses
    from pcgcv1_b200 import transform
    from pcgcv1_b200.models import model_voxception
    vox = TH.CubeFiles.voxels(files[:2])
    out = transform.compress_hyper(vox, model_voxception, a_dir, decompress=True)
    xs = transform.decompress_hyper(*[o.numpy() for o in out[:8]], model_voxception, a_dir)
    assert np.array_equal(xs.numpy(), out[8].numpy())
    # init_ckpt_dir seeds a NEW run: variables restored, global step back at 0
    c_dir = os.path.join(str(tmp_path), "run_c")
    tr_c = TH.train(files, c_dir, num_iteration=0, init_ckpt_dir=a_dir, save_step=2, **kw)
    wc = tr_c.export_weights()
    for k in wa:
        assert np.array_equal(wa[k], wc[k]), k
