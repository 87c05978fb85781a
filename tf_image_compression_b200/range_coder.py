"""Drop-in for the third-party `range_coder` package the reference imports (encode.py:9, decode.py:9):
RangeEncoder / RangeDecoder / prob_to_cum_freq / cum_freq_to_prob with the same call signatures and error
behaviour (other/test_range_coder.py), over the C ABI of librangecoder.so (include/tic_rangecoder.h).

`encode` / `decode` also accept and return numpy uint8 arrays (the codec's native symbol type), which
removes the `astype(int).tolist()` round trip of encode.py:175-182 — seconds per image batch in the
reference — from the path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_rc = None

_SIG = {
    "tic_rc_encoder_open": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p]),
    "tic_rc_encode_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "tic_rc_encode_i32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "tic_rc_encoder_close": (C.c_int, [C.c_void_p]),
    "tic_rc_encoder_free": (None, [C.c_void_p]),
    "tic_rc_encoder_bytes": (C.c_int64, [C.c_void_p]),
    "tic_rc_decoder_open": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p]),
    "tic_rc_decode_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "tic_rc_decode_i32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "tic_rc_decoder_close": (C.c_int, [C.c_void_p]),
    "tic_rc_decoder_free": (None, [C.c_void_p]),
    "tic_rc_prob_to_cum_freq": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]),
    "tic_rc_max_encoded_bytes": (C.c_int64, [C.c_int64]),
    "tic_rc_encode_streams": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int]),
    "tic_rc_decode_streams": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int]),
    "tic_rc_crc32c": (C.c_uint32, [C.c_void_p, C.c_uint64]),
}
ERR_IO, ERR_TABLE, ERR_SYMBOL, ERR_CLOSED = -1, -2, -3, -4


def load():
    global _rc
    if _rc is not None:
        return _rc
    if not L.RC_LIB_PATH.exists():
        raise RuntimeError(f"{L.RC_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(str(L.RC_LIB_PATH))
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _rc = lib
    return lib


def _table(cum_freq):
    """cumFreq -> contiguous uint32 array; OverflowError for entries outside [0, 2^32) like the C++ extension's
    unsigned-int conversion (other/test_range_coder.py:13-34,70-72)."""
    vals = [int(v) for v in cum_freq]
    for v in vals:
        if v < 0 or v >= 2 ** 32:
            raise OverflowError("cumulative frequencies must fit into an unsigned 32-bit integer")
    return np.asarray(vals, dtype=np.uint32)  # totals above 2^16 are a ValueError from the C side (include/tic_rc_core.h)


def _raise(rc, what):
    if rc == ERR_TABLE:
        raise ValueError(f"{what}: invalid cumulative frequency table")
    if rc == ERR_SYMBOL:
        raise ValueError(f"{what}: symbol out of range or of zero probability")
    if rc == ERR_CLOSED:
        raise RuntimeError(f"{what}: file is closed")
    raise RuntimeError(f"{what}: I/O error")


class RangeEncoder:
    def __init__(self, filepath):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.tic_rc_encoder_open(C.byref(self._h), str(filepath).encode())
        if rc != 0:
            self._h = None
            raise RuntimeError(f"cannot open {filepath!r} for writing")

    def encode(self, data, cumFreq):
        if self._h is None:
            raise RuntimeError("encoder was destroyed")
        cum = _table(cumFreq)
        if isinstance(data, np.ndarray) and data.dtype == np.uint8:
            sym = np.ascontiguousarray(data).reshape(-1)
            rc = self._lib.tic_rc_encode_u8(self._h, sym.ctypes.data, sym.size, cum.ctypes.data, cum.size)
        else:
            sym = np.ascontiguousarray(np.asarray(data, dtype=np.int64).reshape(-1))
            if sym.size and (sym.min() < -2 ** 31 or sym.max() >= 2 ** 31):
                raise OverflowError("symbol does not fit into an int")
            sym = sym.astype(np.int32)
            rc = self._lib.tic_rc_encode_i32(self._h, sym.ctypes.data, sym.size, cum.ctypes.data, cum.size)
        if rc != 0:
            _raise(rc, "RangeEncoder.encode")

    def close(self):
        if self._h is not None:
            rc = self._lib.tic_rc_encoder_close(self._h)
            if rc != 0:
                _raise(rc, "RangeEncoder.close")

    @property
    def bytes_written(self):
        return int(self._lib.tic_rc_encoder_bytes(self._h)) if self._h is not None else 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                self._lib.tic_rc_encoder_free(self._h)
                self._h = None
        except Exception:
            pass


class RangeDecoder:
    def __init__(self, filepath):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.tic_rc_decoder_open(C.byref(self._h), str(filepath).encode())
        if rc != 0:
            self._h = None
            raise RuntimeError(f"cannot open {filepath!r} for reading")

    def decode(self, size, cumFreq, dtype=None):
        """`size` symbols as a Python list of ints (the package's contract); dtype=np.uint8 returns an array."""
        if self._h is None:
            raise RuntimeError("decoder was destroyed")
        cum = _table(cumFreq)
        n = int(size)
        if dtype is not None and np.dtype(dtype) == np.uint8:
            out = np.empty(n, dtype=np.uint8)
            rc = self._lib.tic_rc_decode_u8(self._h, out.ctypes.data, n, cum.ctypes.data, cum.size)
            if rc != 0:
                _raise(rc, "RangeDecoder.decode")
            return out
        out = np.empty(n, dtype=np.int32)
        rc = self._lib.tic_rc_decode_i32(self._h, out.ctypes.data, n, cum.ctypes.data, cum.size)
        if rc != 0:
            _raise(rc, "RangeDecoder.decode")
        return out.tolist()

    def close(self):
        if self._h is not None:
            self._lib.tic_rc_decoder_close(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                self._lib.tic_rc_decoder_free(self._h)
                self._h = None
        except Exception:
            pass


def encode_streams(streams, cumFreq, threads=0):
    """Code many whole streams at once on a thread pool (one per image in encode.py's loop, :152-202, "To be
    paralleled"): `streams` is a list of uint8 arrays (or one 2-D uint8 array, one stream per row).  Returns a list of
    `bytes`, each exactly what RangeEncoder(path).encode(stream, cumFreq); close() writes to its file."""
    lib = load()
    cum = _table(cumFreq)
    if isinstance(streams, np.ndarray) and streams.ndim == 2 and streams.dtype == np.uint8:
        sym = np.ascontiguousarray(streams).reshape(-1)
        lens = np.full(streams.shape[0], streams.shape[1], dtype=np.int64)
    else:
        arrs = [np.ascontiguousarray(np.asarray(s, dtype=np.uint8)).reshape(-1) for s in streams]
        lens = np.asarray([a.size for a in arrs], dtype=np.int64)
        sym = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    n = int(lens.size)
    soff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=soff[1:])
    caps = np.asarray([lib.tic_rc_max_encoded_bytes(int(v)) for v in lens], dtype=np.int64)
    ooff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(caps, out=ooff[1:])
    out = np.empty(int(ooff[-1]), dtype=np.uint8)
    nbytes = np.zeros(n, dtype=np.int64)
    rc = lib.tic_rc_encode_streams(sym.ctypes.data, soff.ctypes.data, n, cum.ctypes.data, cum.size, out.ctypes.data,
                                   ooff.ctypes.data, nbytes.ctypes.data, int(threads))
    if rc != 0:
        _raise(rc, "encode_streams")
    return [out[int(ooff[i]):int(ooff[i] + nbytes[i])].tobytes() for i in range(n)]


def decode_streams(blobs, sizes, cumFreq, threads=0):
    """Decode many whole streams at once (decode.py:171-208 over a directory): blobs[i] holds sizes[i] symbols.
    Returns a list of uint8 arrays."""
    lib = load()
    cum = _table(cumFreq)
    n = len(blobs)
    lens = np.asarray([len(b) for b in blobs], dtype=np.int64)
    ioff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=ioff[1:])
    data = np.frombuffer(b"".join(bytes(b) for b in blobs), dtype=np.uint8) if n else np.zeros(0, np.uint8)
    if data.size == 0:
        data = np.zeros(1, np.uint8)
    sizes = np.asarray(sizes, dtype=np.int64)
    soff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes, out=soff[1:])
    sym = np.empty(int(soff[-1]), dtype=np.uint8)
    rc = lib.tic_rc_decode_streams(data.ctypes.data, ioff.ctypes.data, lens.ctypes.data, n, cum.ctypes.data, cum.size,
                                   sym.ctypes.data, soff.ctypes.data, int(threads))
    if rc != 0:
        _raise(rc, "decode_streams")
    return [sym[int(soff[i]):int(soff[i + 1])] for i in range(n)]


def crc32c(data) -> int:
    """CRC-32C (Castagnoli) of a bytes-like object (tic_rc_crc32c, host-only library)."""
    b = bytes(data)
    return int(load().tic_rc_crc32c(b, len(b)))


def prob_to_cum_freq(prob, resolution=1024):
    """Cumulative frequency table (list of n + 1 ints ending at `resolution`) for a probability vector:
    non-zero probabilities get non-zero width, zero probabilities zero width (other/test_range_coder.py:186-229)."""
    p = np.ascontiguousarray(np.asarray(prob, dtype=np.float64).reshape(-1))
    cum = np.zeros(p.size + 1, dtype=np.uint32)
    rc = load().tic_rc_prob_to_cum_freq(p.ctypes.data, p.size, int(resolution), cum.ctypes.data)
    if rc != 0:
        raise ValueError("invalid probabilities / resolution")
    return [int(v) for v in cum]


def cum_freq_to_prob(cumFreq):
    c = np.asarray(cumFreq, dtype=np.float64)
    return np.diff(c) / c[-1]
