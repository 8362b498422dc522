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
