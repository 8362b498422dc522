"""Far-field tiles of the analysis transform: with PCGC_FARFIELD=1 the tile kernel copies every tile of the three VRN-16
blocks whose receptive field holds no occupied voxel from the cached activations of the all-zero cube (the default; PCGC_FARFIELD=0
computes every tile).  The latents must be the same
BITS as with every tile computed -- for cubes of the workload, the empty cube, a cube whose every tile is near-field, and single
voxels at a corner / an edge / the centre (SAME padding at the cube faces is part of the empty cube's activations)."""
import numpy as np
import pytest
import torch

from pcgcv1_b200 import runtime, synthetic

pytestmark = pytest.mark.gpu


def _cases():
    cubes, _, _ = synthetic.workload("vox10", seed=0, max_cubes=70)           # more than one sub-batch of 64
    extra = np.zeros((6, 64, 64, 64, 1), np.uint8)
    rng = np.random.default_rng(3)
    extra[1][rng.random((64, 64, 64, 1)) < 0.02] = 1                           # scattered: (almost) every tile is near-field
    extra[2][0, 0, 0, 0] = 1                                                   # corner
    extra[3][63, 31, 0, 0] = 1                                                 # edge
    extra[4][32, 32, 32, 0] = 1                                                # centre
    extra[5][:, 40, :, 0] = 1                                                  # a full plane
    return np.concatenate([cubes, extra])


def test_far_field_tiles_give_the_same_bits(codec, monkeypatch):
    x = codec.to_device(_cases())
    monkeypatch.setenv("PCGC_FARFIELD", "0")
    ref = codec.analysis(x).clone()
    monkeypatch.setenv("PCGC_FARFIELD", "1")
    got = codec.analysis(x).clone()
    again = codec.analysis(x[:7].contiguous()).clone()                          # another batch size, the cached empty-cube activations reused
    codec.synchronize()
    assert torch.equal(ref, got)
    assert torch.equal(ref[:7], again)
    # the compress path end to end: same stream
    from pcgcv1_b200 import transform
    from pcgcv1_b200.models import model_voxception
    cubes = _cases()[:20]
    monkeypatch.setenv("PCGC_FARFIELD", "0")
    a = [o.numpy() for o in transform.compress_hyper(cubes, model_voxception, "")]
    monkeypatch.setenv("PCGC_FARFIELD", "1")
    b = [o.numpy() for o in transform.compress_hyper(cubes, model_voxception, "")]
    for u, v in zip(a, b):
        assert np.array_equal(np.asarray(u), np.asarray(v))
