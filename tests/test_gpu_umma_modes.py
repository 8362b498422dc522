"""The conv engine has two kernels (tile and z-streaming, csrc/umma_conv.cu) selected per shape; PCGC_UMMA_STREAM is read once
per process, so the other two settings run the shape tests in a subprocess: 0 = tile kernel everywhere, 2 = stream every shape
the streaming kernel covers (including the Cin = 8 paired-tap form that defaults to the tile kernel)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["0", "2"])
def test_umma_shapes_in_other_stream_modes(mode):
    env = dict(os.environ, PCGC_UMMA_STREAM=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_umma.py"), "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
