"""The conv engine has three kernels (tile, z-streaming, z-banded; csrc/umma_conv.cu) selected per shape.  PCGC_UMMA_STREAM and
PCGC_UMMA_ZBAND are read once per process, so the other kernel-selection settings run the shape tests in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,zband,kb", [("0", "1", "0"), ("2", "1", "0"), ("1", "0", "0"), ("1", "1", "1")])
def test_umma_shapes_in_other_stream_modes(mode, zband, kb):
    """(stream, z-band): (0, -) tile kernel everywhere; (2, 1) every streaming form incl. the paired-tap one; (1, 0) the plain
    streaming kernel where the default build uses the z-banded one; kb = 1 adds the opt-in 32-column z-banded forms."""
    env = dict(os.environ, PCGC_UMMA_STREAM=mode, PCGC_UMMA_ZBAND=zband, PCGC_KB_ZBAND=kb)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_umma.py"), "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


_HD_SCRIPT = r"""
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, %r)
from pcgcv1_b200 import _lib, runtime
codec = runtime.get_codec("voxception", "")
if int(sys.argv[1]):
    codec.set_engine(_lib.ENGINE_FFMA)
rng = np.random.default_rng(7)
z = torch.from_numpy(rng.integers(-6, 7, size=(5, 8, 8, 8, 8)).astype(np.float32)).to(codec.dev)
loc, scale = codec.hyper_decode(z, 1e-9)
loc1, scale1 = codec.hyper_decode(z[3:4], 1e-9)
codec.synchronize()
assert torch.equal(loc[3:4], loc1) and torch.equal(scale[3:4], scale1)
print("HD", hashlib.sha256(loc.cpu().numpy().tobytes() + scale.cpu().numpy().tobytes()).hexdigest())
"""


def test_hyper_decoder_bits_do_not_depend_on_tuning_switches_or_engine():
    """loc / scale become integer CDF tables on both sides of a stream: every kernel-selection switch and the engine setting
    must give the SAME bits (the hyper decoder runs one pinned program)."""
    settings = [({}, 0), ({"PCGC_UMMA_STREAM": "0"}, 0), ({"PCGC_UMMA_STREAM": "2", "PCGC_KB_ZBAND": "1"}, 0), ({"PCGC_UMMA_ZBAND": "0"}, 0),
                ({"PCGC_UMMA_ZT": "2", "PCGC_UMMA_WT": "1"}, 0), ({"PCGC_SUB_BATCH": "2"}, 0), ({}, 1)]
    digests = []
    for env_add, ffma in settings:
        env = dict(os.environ, **env_add)
        r = subprocess.run([sys.executable, "-c", _HD_SCRIPT % ROOT, str(ffma)], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        digests.append([l for l in r.stdout.splitlines() if l.startswith("HD ")][-1])
    assert len(set(digests)) == 1, digests
