"""TF-1.13 checkpoint import.  Parity unpinned: no checkpoint of the reference exists offline,
so these tests pin the TensorBundle reader against this repo's own writer of the same published layout (prefix-compressed
blocks, several data blocks, checksums) and the attribute-path -> Keras-name mapping against the reference's model source."""
import os
import struct

import numpy as np
import pytest

from pcgcv1_b200 import tf_checkpoint as tfc
from pcgcv1_b200 import weights as W


def test_crc32c_known_answers():
    assert tfc.crc32c(b"123456789") == 0xE3069283                       # the standard CRC-32C check value
    assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA                         # RFC 3720 test vector
    assert tfc.masked_crc(b"") == 0xA282EAD8


@pytest.mark.parametrize("model", ["voxception", "simple"])
def test_export_import_round_trip(tmp_path, model):
    w = W.synthetic_weights(model)
    w = {k: v for k, v in w.items() if not k.startswith("estimator_y/")}
    prefix = tfc.export_checkpoint(str(tmp_path), w, step=7)
    assert os.path.exists(prefix + ".index") and tfc.latest_checkpoint(str(tmp_path)) == prefix
    raw = tfc.read_bundle(prefix, verify_tensors=True)
    assert "synthesis_transform/vrn1_1/conv1_1/kernel" + tfc.SUFFIX in raw or model == "simple"
    got = W.load(str(tmp_path), model)                                     # weights.npz absent -> TF checkpoint path
    assert sorted(got) == sorted(w)
    for k in w:
        assert got[k].dtype == w[k].dtype and got[k].shape == w[k].shape and np.array_equal(got[k], w[k]), k


def test_name_mapping_follows_the_reference_attributes():
    t = {n + tfc.SUFFIX: np.zeros(1, np.float32) for n in [
        "analysis_transform/conv_in/kernel", "analysis_transform/vrn2_3/conv2_2/bias", "synthesis_transform/vrn3_1/conv1_2/kernel",
        "synthesis_transform/up_1/bias", "hyper_encoder/conv2/kernel", "hyper_decoder/conv2/kernel", "hyper_decoder/conv4_2/bias",
        "estimator/matrix_0", "estimator/bais_3", "estimator/_factors/2"]}
    t["global_step" + tfc.SUFFIX] = np.zeros(1, np.int64)
    t["main_optimizer/beta1_power" + tfc.SUFFIX] = np.zeros(1, np.float32)
    t["analysis_transform/conv_in/kernel/.OPTIMIZER_SLOT/main_optimizer/m" + tfc.SUFFIX] = np.zeros(1, np.float32)
    t["_CHECKPOINTABLE_OBJECT_GRAPH"] = np.zeros(1, np.uint8)
    got = sorted(tfc.map_names(t))
    assert got == sorted(["analysis_transform/conv_in/kernel", "analysis_transform/vrn2_3_conv2_2/bias", "synthesis_transform/dvrn3_1_conv1_2/kernel",
                          "synthesis_transform/up_1/bias", "hyper_encoder/conv2/kernel", "hyper_decoder/deconv2/kernel",
                          "hyper_decoder/deconv4_2/bias", "estimator/matrix_0", "estimator/bais_3", "estimator/factor_2"])
    with pytest.raises(KeyError):
        tfc.map_names({"analysis_transform/mystery/gamma" + tfc.SUFFIX: np.zeros(1, np.float32)})


def test_corrupt_index_is_detected(tmp_path):
    prefix = str(tmp_path / "ckpt-1")
    tfc.write_bundle(prefix, {"a/kernel" + tfc.SUFFIX: np.arange(6, dtype=np.float32).reshape(2, 3)})
    data = bytearray(open(prefix + ".index", "rb").read())
    data[3] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(data))
    with pytest.raises(ValueError):
        tfc.read_bundle(prefix)
    open(prefix + ".index", "wb").write(bytes(data[:-8]) + struct.pack("<Q", 1))
    with pytest.raises(ValueError):
        tfc.read_bundle(prefix)
    with pytest.raises(FileNotFoundError):
        W.load(str(tmp_path / "nothing_here"), "voxception")
