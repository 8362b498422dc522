import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the oracle's C restatement and the product library once per session (both are
    compile-only steps; nvcc cross-compiles without a GPU)."""
    from oracle import build as obuild
    obuild.build()
    from pcgcv1_b200 import build as pbuild
    pbuild.build()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        with np.load(os.path.join(GOLDEN, name)) as z:
            return {k: z[k] for k in z.files}
    return load


@pytest.fixture(scope="session")
def codec():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pcgcv1_b200 import runtime
    return runtime.get_codec("voxception", "")


@pytest.fixture(scope="session")
def codec_simple():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pcgcv1_b200 import runtime
    return runtime.get_codec("simple", "")
