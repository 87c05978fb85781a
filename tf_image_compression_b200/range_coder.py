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
    return np.asarray(vals, dtype=np.uint32)


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
