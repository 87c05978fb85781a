"""TF-V2 checkpoint ("tensor bundle") reader / writer without TensorFlow — what tf.train.Saver().restore(sess,
'model_N/params_for_test/params') reads in utils.restore_params (utils/utils.py:84-93), so the reference's params
layout (params.index + params.data-00000-of-00001 + checkpoint) keeps working when trained weights are supplied.

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*, restated from the published layout;
PARITY UNPINNED: the reference ships no checkpoint and TensorFlow is not installable here, so the reader is
checked against this module's own writer and hand-assembled blocks, not against a TF-written file):

* `<prefix>.index` is a LevelDB-style sorted string table.  Footer (last 48 bytes): metaindex BlockHandle, index
  BlockHandle (varint64 offset, varint64 size), zero padding to 40 bytes, magic 0xdb4775248b80fb57 (LE).  A block is
  `entries | uint32 restarts[n] | uint32 n` followed by a 5-byte trailer (compression type, masked crc32c of
  contents + type).  An entry is varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value.
  The index block maps separator keys to data-block handles.
* key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: VersionDef}; every other key is a variable name ->
  BundleEntryProto {1: dtype, 2: TensorShapeProto {2: dim {1: size}}, 3: shard_id, 4: offset, 5: size,
  6: fixed32 masked crc32c of the bytes, 7: slices (partitioned variables: not used by this codebase)}.
* `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at [offset, offset + size).

Variable names of this codebase: '<scope>/kernel', '<scope>/bias' with scope = 'encode_0', 'encode_res_1/conv_0', …
(basic_block/basic_block.py:27-71: tf.variable_scope(name) + tf.get_variable('kernel' | 'bias')); optimizer slots
('…/Adam', 'beta1_power') that tf.train.Saver() also stores are returned like any other tensor and ignored by
Codec.load_params."""
from __future__ import annotations

import os
import struct
from pathlib import Path

import numpy as np


MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           19: np.float16, 17: np.uint16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(ValueError):
    pass


def crc32c(data) -> int:
    """CRC-32C (Castagnoli) of any bytes-like object through the host-only library (tic_rc_crc32c in
    librangecoder.so: reading a checkpoint never needs the CUDA library)."""
    from . import range_coder
    return range_coder.crc32c(data)


def masked_crc32c(data) -> int:
    """crc32c::Mask: rotate right by 15 and add a constant (stored CRCs of data that itself embeds CRCs)."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / protobuf wire format ---------------------------------------------------------------
def _get_varint(buf, pos):
    shift = 0
    val = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """-> {field: [values]}; varint fields as int, fixed32/64 as int, length-delimited as bytes."""
    out = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            if pos + n > len(buf):
                raise CheckpointError("truncated length-delimited field")
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _field(tag, wt, payload):
    return _put_varint((tag << 3) | wt) + payload


def _msg(tag, body):
    return _field(tag, 2, _put_varint(len(body)) + body)


# ---- snappy (raw format) decoder: TF's bundle writer stores the index uncompressed, other table writers may not ----
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("corrupt snappy block")
        for _ in range(ln):  # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("snappy length mismatch")
    return bytes(out)


# ---- table reader ------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    if offset + size + 5 > len(data):
        raise CheckpointError("block handle points outside the index file")
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if masked_crc32c(data[offset:offset + size + 1]) != stored:
            raise CheckpointError(f"index block at {offset}: crc32c mismatch")
    if ctype == 1:
        raw = _snappy_decompress(raw)
    elif ctype != 0:
        raise CheckpointError(f"unknown block compression type {ctype}")
    return raw


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("block too small")
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    if end < 0:
        raise CheckpointError("bad restart array")
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise CheckpointError("corrupt block entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(prefix, verify=True):
    """-> (header dict, {name: entry dict}) of `<prefix>.index`."""
    path = str(prefix) + ".index"
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != MAGIC:
        raise CheckpointError(f"{path}: not a TF-V2 checkpoint index (bad table magic)")
    footer = data[len(data) - 48:]
    pos = 0
    _, pos = _get_varint(footer, pos)      # metaindex offset
    _, pos = _get_varint(footer, pos)      # metaindex size
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    entries = {}
    header = None
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p2 = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p2)
        for key, value in _block_entries(_read_block(data, boff, bsize, verify)):
            msg = _parse_proto(value)
            if key == b"":
                header = {"num_shards": msg.get(1, [1])[0], "endianness": msg.get(2, [0])[0]}
                continue
            shape = []
            if 2 in msg:
                for dim in _parse_proto(msg[2][0]).get(2, []):
                    size = _parse_proto(dim).get(1, [0])[0]
                    shape.append(size - (1 << 64) if size >> 63 else size)
            entries[key.decode("utf-8")] = {
                "dtype": msg.get(1, [0])[0], "shape": tuple(shape), "shard_id": msg.get(3, [0])[0],
                "offset": msg.get(4, [0])[0], "size": msg.get(5, [0])[0], "crc32c": msg.get(6, [None])[0],
                "sliced": 7 in msg}
    if header is None:
        raise CheckpointError(f"{path}: no bundle header entry")
    if header["endianness"] != 0:
        raise CheckpointError("big-endian bundles are not supported")
    return header, entries


def read_checkpoint(prefix, names=None, verify=True):
    """{variable name: ndarray} of the TF-V2 checkpoint `<prefix>` (all variables, or `names`)."""
    header, entries = read_index(prefix, verify)
    shards = {}
    out = {}
    for name in (entries if names is None else names):
        if name not in entries:
            raise KeyError(f"checkpoint {prefix} has no variable {name!r}")
        e = entries[name]
        if e["sliced"]:
            raise CheckpointError(f"{name}: partitioned variables are not supported")
        if e["dtype"] not in _DTYPES:
            raise CheckpointError(f"{name}: unsupported dtype enum {e['dtype']}")
        dt = np.dtype(_DTYPES[e["dtype"]])
        sid = e["shard_id"]
        if sid not in shards:
            p = f"{prefix}.data-{sid:05d}-of-{header['num_shards']:05d}"
            shards[sid] = np.memmap(p, dtype=np.uint8, mode="r") if os.path.getsize(p) else np.zeros(0, np.uint8)
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if raw.size != e["size"] or count * dt.itemsize != e["size"]:
            raise CheckpointError(f"{name}: {e['size']} bytes on record, shape {e['shape']} {dt} needs {count * dt.itemsize}")
        rb = raw.tobytes()
        if verify and e["crc32c"] is not None and masked_crc32c(rb) != e["crc32c"]:
            raise CheckpointError(f"{name}: crc32c mismatch in the data shard")
        out[name] = np.frombuffer(rb, dtype=dt).reshape(e["shape"]).copy()
    return out


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file, else None."""
    state = Path(directory) / "checkpoint"
    if not state.exists():
        return None
    for line in state.read_text().splitlines():
        if line.startswith("model_checkpoint_path:"):
            name = line.split(":", 1)[1].strip().strip('"')
            p = Path(name)
            return str(p if p.is_absolute() else Path(directory) / p)
    return None


# ---- writer (saving seeded / converted parameters in the reference's layout; also the reader's test vector) --------
def _build_block(items, restart_interval):
    body = bytearray()
    restarts = []
    prev = b""
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        prev = key
    if not restarts:
        restarts.append(0)
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def write_checkpoint(prefix, tensors, block_size=4096):
    """Write {name: ndarray} as `<prefix>.index`, `<prefix>.data-00000-of-00001` and a `checkpoint` state file."""
    prefix = str(prefix)
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    items = [(b"", _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) + _msg(3, _field(1, 0, _put_varint(1))))]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])  # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in _DTYPE_IDS:
                raise CheckpointError(f"{name}: dtype {a.dtype} has no TF enum here")
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes(order="C")
            shape = b"".join(_msg(2, _field(1, 0, _put_varint(int(d)))) for d in a.shape)
            entry = _field(1, 0, _put_varint(_DTYPE_IDS[a.dtype])) + _msg(2, shape)
            if offset:
                entry += _field(4, 0, _put_varint(offset))
            entry += _field(5, 0, _put_varint(len(raw))) + _field(6, 5, struct.pack("<I", masked_crc32c(raw)))
            items.append((name.encode("utf-8"), entry))
            f.write(raw)
            offset += len(raw)
    out = bytearray()

    def emit(block):
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block)
        out.append(0)  # kNoCompression, like BundleWriter
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return handle

    index_items = []
    cur = []
    cur_bytes = 0
    for key, value in items:
        cur.append((key, value))
        cur_bytes += len(key) + len(value) + 3
        if cur_bytes >= block_size:
            index_items.append((cur[-1][0], emit(_build_block(cur, 16))))
            cur, cur_bytes = [], 0
    if cur:
        index_items.append((cur[-1][0], emit(_build_block(cur, 16))))
    meta = emit(_build_block([], 1))
    index = emit(_build_block(index_items, 1))
    footer = meta + index
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')


def restore_params(codec, params_file=None, model_num=0, root=".", verify=True):
    """utils.restore_params (utils/utils.py:84-93): load `<root>/model_<N>/params_for_test/params` (or the explicit
    `params_file` prefix, the reference's -p flag) into the codec's encoder and decoder graphs."""
    from . import _lib
    if not params_file:
        params_file = str(Path(root) / f"model_{model_num}" / "params_for_test" / "params")
    need = [l.scope + s for l in list(codec.enc_layers) + list(codec.dec_layers) for s in ("/kernel", "/bias")]
    params = read_checkpoint(params_file, names=need, verify=verify)
    codec.load_params(_lib.GRAPH_ENCODER, codec.enc_layers, params)
    codec.load_params(_lib.GRAPH_DECODER, codec.dec_layers, params)
    return params_file
