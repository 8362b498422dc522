"""Drop-in for ``myutils/gpcc_wrapper.py``: ``gpcc_encode(filedir, bin_dir)`` / ``gpcc_decode(bin_dir, rec_dir)`` code the
cube-position list (a tiny .ply, ~90 bytes of side information per cloud) losslessly.

The reference shells out to a prebuilt MPEG TMC13 v6 binary, ``myutils/tmc3`` relative to the working directory
(gpcc_wrapper.py:5-42); no source for it exists in the reference, so it is not rebuilt here.  When that binary is present
it is called with the reference's exact command lines, so ``.cubepos`` files interoperate.  When it is absent the positions
are stored raw (magic ``PCGCRAW1`` + uint8 triples: 3 bytes per cube) and ``gpcc_decode`` recognises the magic."""
from __future__ import annotations

import os
import subprocess

import numpy as np

TMC3 = "myutils/tmc3"
RAW_MAGIC = b"PCGCRAW1"


def have_tmc3() -> bool:
    return os.path.isfile(TMC3) and os.access(TMC3, os.X_OK)


def _run(cmd, show):
    subp = subprocess.Popen(cmd, shell=True, stdout=subprocess.PIPE)
    c = subp.stdout.readline()
    while c:
        if show:
            print(c)
        c = subp.stdout.readline()
    subp.wait()


def gpcc_encode(filedir, bin_dir, show=False):
    """Cube positions .ply -> compressed stream (gpcc_wrapper.py:5-28)."""
    if have_tmc3():
        _run(TMC3 + ' --mode=0' + ' --positionQuantizationScale=1' + ' --trisoup_node_size_log2=0' +
             ' --ctxOccupancyReductionFactor=3' + ' --neighbourAvailBoundaryLog2=8' + ' --intra_pred_max_node_size_log2=6' +
             ' --inferredDirectCodingMode=0' + ' --uncompressedDataPath=' + filedir + ' --compressedStreamPath=' + bin_dir, show)
        return
    from ..dataprocess.inout_points import load_ply_data
    pos = load_ply_data(filedir)
    with open(bin_dir, "wb") as f:
        f.write(RAW_MAGIC)
        f.write(np.asarray(pos, dtype=np.uint8).tobytes())


def gpcc_decode(bin_dir, rec_dir, show=False):
    """Compressed stream -> cube positions .ply (gpcc_wrapper.py:30-42)."""
    with open(bin_dir, "rb") as f:
        head = f.read(len(RAW_MAGIC))
        body = f.read() if head == RAW_MAGIC else None
    if body is not None:
        from ..dataprocess.inout_points import write_ply_data
        write_ply_data(rec_dir, np.frombuffer(body, dtype=np.uint8).reshape(-1, 3))
        return
    if not have_tmc3():
        raise FileNotFoundError("%s is a G-PCC stream but %s is not present (the reference ships it prebuilt)" % (bin_dir, TMC3))
    _run(TMC3 + ' --mode=1' + ' --compressedStreamPath=' + bin_dir + ' --reconstructedDataPath=' + rec_dir, show)
