"""The C-ABI library loads and exports every symbol include/pcgc_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

from pcgcv1_b200 import _lib, netspec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pcgc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcgc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libpcgc_b200.so does not export %s" % n
    # and the ctypes stub binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_enums():
    L = _lib.lib()
    assert L.pcgc_abi_version() == 1
    hdr = open(os.path.join(ROOT, "include", "pcgc_b200.h")).read()
    assert "PCGC_MAX_SYMBOLS %d" % _lib.MAX_SYMBOLS in hdr
    for name, val in (("PCGC_NET_VOX_ANALYSIS", 0), ("PCGC_NET_SIMPLE_SYNTHESIS", 5), ("PCGC_ERR_BAD_RANGE", -2)):
        assert re.search(r"%s = %d\b" % (name, val), hdr)


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    L = _lib.lib()
    h = C.c_void_p()
    assert L.pcgc_create(C.byref(h), 0) != 0 and not h.value
    import pytest
    from pcgcv1_b200 import runtime
    with pytest.raises(RuntimeError):
        runtime.Codec("voxception", "")


def test_mac_counts_match_survey():
    """algorithmic GFLOP per cube (roofline numerators)."""
    a = netspec.macs_per_cube(netspec.NETS[("voxception", "analysis_transform")], 64)
    s = netspec.macs_per_cube(netspec.NETS[("voxception", "synthesis_transform")], 16)
    he = netspec.macs_per_cube(netspec.NETS[("voxception", "hyper_encoder")], 16)
    hd = netspec.macs_per_cube(netspec.NETS[("voxception", "hyper_decoder")], 8)
    assert abs(2 * a / 1e9 - 10.3998) < 1e-3 and abs(2 * s / 1e9 - 10.3998) < 1e-3
    assert abs(2 * he / 1e9 - 0.0672) < 1e-3 and abs(2 * hd / 1e9 - 0.3504) < 1e-3
    sa = netspec.macs_per_cube(netspec.NETS[("simple", "analysis_transform")], 64)
    ss = netspec.macs_per_cube(netspec.NETS[("simple", "synthesis_transform")], 8)
    assert abs(2 * (sa + ss) / 1e9 - 5.4169) < 1e-3
