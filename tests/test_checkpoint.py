"""TF-V2 checkpoint bundle reader / writer (utils.restore_params, utils/utils.py:84-93) — CPU tests.
Parity unpinned against TensorFlow (no TF, no checkpoint in the reference): the reader is checked against a
hand-assembled index, against the module's own writer, and on corrupted files."""
import struct

import numpy as np
import pytest

from tf_image_compression_b200 import checkpoint as K
from tf_image_compression_b200 import variants as V
from tf_image_compression_b200.codec import reference_init


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors + the classic check value
    assert K.crc32c(b"123456789") == 0xE3069283
    assert K.crc32c(bytes(32)) == 0x8A9136AA
    assert K.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert K.crc32c(bytes(range(32))) == 0x46DD794E
    assert K.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert K.crc32c(b"") == 0
    c = K.crc32c(b"foo")
    assert K.masked_crc32c(b"foo") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _block(entries_bytes, restarts):
    return entries_bytes + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def _with_trailer(block):
    return block + b"\x00" + struct.pack("<I", K.masked_crc32c(block + b"\x00"))


def test_reads_a_hand_assembled_bundle(tmp_path):
    """Every byte of the index spelt out: header entry, one float32[2] variable with a shared key prefix."""
    values = np.array([1.5, -2.25], np.float32)
    raw = values.tobytes()
    (tmp_path / "params.data-00000-of-00001").write_bytes(b"\xAA" * 4 + raw)  # tensor at offset 4
    header = bytes([0x08, 0x01, 0x10, 0x00, 0x1A, 0x02, 0x08, 0x01])  # num_shards=1, LITTLE, version{producer=1}
    shape = bytes([0x12, 0x02, 0x08, 0x02])  # dim { size: 2 }
    entry = bytes([0x08, 0x01, 0x12, len(shape)]) + shape + bytes([0x20, 0x04, 0x28, 0x08, 0x35]) + \
        struct.pack("<I", K.masked_crc32c(raw))  # dtype DT_FLOAT, shape, offset 4, size 8, fixed32 crc
    entry2 = bytes([0x08, 0x03, 0x12, 0x00, 0x28, 0x04, 0x35]) + struct.pack("<I", K.masked_crc32c(b"\xAA" * 4))  # int32 scalar
    e0 = bytes([0, 0, len(header)]) + header                                  # key ""
    e1 = bytes([0, 6, len(entry)]) + b"a/bias" + entry                        # key "a/bias"
    e2 = bytes([2, 4, len(entry2)]) + b"step" + entry2                        # shares "a/" -> key "a/step"
    data_block = _block(e0 + e1 + e2, [0])
    out = bytearray(_with_trailer(data_block))
    meta_off = len(out)
    meta_block = _block(b"", [0])
    out += _with_trailer(meta_block)
    idx_off = len(out)
    handle = bytes([0, len(data_block)])
    idx_block = _block(bytes([0, 6, len(handle)]) + b"a/step" + handle, [0])
    out += _with_trailer(idx_block)
    footer = bytes([meta_off, len(meta_block), idx_off, len(idx_block)])
    out += footer + bytes(40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    (tmp_path / "params.index").write_bytes(bytes(out))
    got = K.read_checkpoint(tmp_path / "params")
    assert set(got) == {"a/bias", "a/step"}
    assert got["a/bias"].dtype == np.float32 and np.array_equal(got["a/bias"], values)
    assert got["a/step"].shape == () and got["a/step"].dtype == np.int32 and int(got["a/step"]) == -1431655766  # 0xAAAAAAAA
    assert np.array_equal(K.read_checkpoint(tmp_path / "params", names=["a/bias"])["a/bias"], values)


@pytest.mark.parametrize("variant", ["model_0", "base_model/ch_128"])
def test_writer_reader_round_trip_in_the_reference_layout(tmp_path, variant):
    """model_N/params_for_test/params{.index,.data-00000-of-00001} + checkpoint, several data blocks, prefix-compressed
    keys, optimizer slots next to the variables."""
    layers = V.encoder_layers(variant) + V.decoder_layers(variant)
    params = reference_init(layers, 99)
    params["beta1_power"] = np.float32(0.9)
    params["encode_res_1/conv_0/kernel/Adam"] = np.zeros((3, 3, 4, 4), np.float32)
    params["global_step"] = np.int64(123456789012)
    prefix = tmp_path / "model_0" / "params_for_test" / "params"
    K.write_checkpoint(prefix, params, block_size=256)
    assert K.latest_checkpoint(prefix.parent) == str(prefix)
    header, entries = K.read_index(prefix)
    assert header["num_shards"] == 1 and list(entries) == sorted(entries, key=lambda s: s.encode())
    got = K.read_checkpoint(prefix)
    assert set(got) == set(params)
    for k, v in params.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape and np.array_equal(got[k], v), k
    first = layers[0].scope + "/kernel"
    assert got[first].shape == layers[0].kernel_shape


def test_corruption_and_missing_names_fail_loudly(tmp_path):
    prefix = tmp_path / "params"
    K.write_checkpoint(prefix, {"w/kernel": np.arange(36, dtype=np.float32).reshape(3, 3, 2, 2), "w/bias": np.zeros(2, np.float32)})
    with pytest.raises(KeyError):
        K.read_checkpoint(prefix, names=["nope/kernel"])
    data = bytearray((tmp_path / "params.data-00000-of-00001").read_bytes())
    data[5] ^= 0x40
    (tmp_path / "params.data-00000-of-00001").write_bytes(bytes(data))
    with pytest.raises(K.CheckpointError, match="crc32c"):
        K.read_checkpoint(prefix)
    assert K.read_checkpoint(prefix, verify=False)["w/kernel"].shape == (3, 3, 2, 2)
    idx = bytearray((tmp_path / "params.index").read_bytes())
    bad = bytearray(idx)
    bad[3] ^= 0x01
    (tmp_path / "params.index").write_bytes(bytes(bad))
    with pytest.raises(K.CheckpointError):
        K.read_checkpoint(prefix)
    idx[-1] ^= 0xFF
    (tmp_path / "params.index").write_bytes(bytes(idx))
    with pytest.raises(K.CheckpointError, match="magic"):
        K.read_checkpoint(prefix)


def test_snappy_block_decoder():
    # "abcabcabcabcX": literal "abc", copy(offset 3, length 9) with a 1-byte offset tag, literal "X"
    comp = bytes([13, (3 - 1) << 2]) + b"abc" + bytes([((9 - 4) << 2) | 1, 3]) + bytes([0 << 2]) + b"X"
    assert K._snappy_decompress(comp) == b"abcabcabcabcX"
    # 2-byte offset copy
    comp2 = bytes([8, (4 - 1) << 2]) + b"wxyz" + bytes([((4 - 1) << 2) | 2, 4, 0])
    assert K._snappy_decompress(comp2) == b"wxyzwxyz"
