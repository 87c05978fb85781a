"""Host entropy coder (tf_image_compression_b200.range_coder over librangecoder.so).

The only golden vector the reference holds for this step is the test-suite of the third-party `range_coder`
package it vendors as other/test_range_coder.py.  Where /root/reference is mounted that file
is run UNMODIFIED against this module (aliased as `range_coder`); its known-answer vector is restated here so
it also runs on the GPU box."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tf_image_compression_b200 import range_coder as RC

ROOT = Path(__file__).resolve().parents[1]
REF_TEST = Path("/root/reference/other/test_range_coder.py")


def test_known_answer_vector(tmp_path):
    """other/test_range_coder.py:37-68: cumFreq [0,4,6,8], 17 x [0,0,0,0,1,2] -> 17 bytes; bytes 4..16 == 0x0b."""
    path = tmp_path / "kat.bin"
    enc = RC.RangeEncoder(str(path))
    enc.encode([0, 0, 0, 0, 1, 2] * 17, [0, 4, 6, 8])
    enc.close()
    raw = path.read_bytes()
    assert len(raw) == 17
    assert raw[4:] == b"\x0b" * 13
    with pytest.raises(RuntimeError):
        enc.encode([0], [0, 4, 6, 8])  # closed
    dec = RC.RangeDecoder(str(path))
    assert dec.decode(6 * 17, [0, 4, 6, 8]) == [0, 0, 0, 0, 1, 2] * 17
    dec.close()


def test_error_conventions(tmp_path):
    enc = RC.RangeEncoder(str(tmp_path / "e.bin"))
    data = [0, 1, 2]
    with pytest.raises(OverflowError):
        enc.encode(data, [-1, 1])
    with pytest.raises(OverflowError):
        enc.encode(data, [0, 1, 2 ** 32])
    for bad in ([1, 2, 3], [0, 1], [0, 8, 8, 8], [], [0], [0, 5, 3, 8]):
        with pytest.raises(ValueError):
            enc.encode(data, bad)
    enc.close()
    dec = RC.RangeDecoder(str(tmp_path / "e.bin"))
    for bad in ([], [0]):
        with pytest.raises(ValueError):
            dec.decode(3, bad)
    assert dec.decode(0, [0, 4, 6, 8]) == []
    with pytest.raises(RuntimeError):
        RC.RangeEncoder(str(tmp_path / "no_such_dir" / "x.bin"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_round_trip_mixed_tables_and_dtypes(tmp_path, seed):
    rs = np.random.RandomState(seed)
    path = str(tmp_path / "rt.bin")
    tables, chunks = [], []
    for _ in range(5):
        k = rs.randint(1, 40)
        cum = [0] + [int(v) for v in np.cumsum(rs.randint(1, 200, size=k))]
        n = rs.randint(0, 3000)
        tables.append(cum)
        chunks.append(rs.randint(0, k, size=n).astype(np.uint8))
    enc = RC.RangeEncoder(path)
    for cum, c in zip(tables, chunks):
        enc.encode(c if seed % 2 else c.tolist(), cum)  # uint8 array fast path and the list path give the same stream
    enc.close()
    assert enc.bytes_written == os.path.getsize(path)
    dec = RC.RangeDecoder(path)
    for cum, c in zip(tables, chunks):
        got = dec.decode(len(c), cum, dtype=np.uint8)
        assert np.array_equal(got, c)
    dec.close()


def test_binary_symbols_reach_the_entropy(tmp_path):
    """q = 2 (every shipped config): 1e6 symbols with p(1) = 0.1 -> within 0.5 % of the table's cross entropy."""
    rs = np.random.RandomState(5)
    sym = (rs.rand(1_000_000) < 0.1).astype(np.uint8)
    prob = np.bincount(sym, minlength=2) / sym.size
    freq = prob * 4096 + 1  # encode.py:82-91
    cum = RC.prob_to_cum_freq(freq / freq.sum(), resolution=4096)
    path = str(tmp_path / "b.bin")
    enc = RC.RangeEncoder(path)
    enc.encode(sym, cum)
    enc.close()
    q = np.diff(cum) / cum[-1]
    ideal_bits = -(np.log2(q[sym.astype(int)])).sum()
    assert os.path.getsize(path) * 8 <= ideal_bits * 1.005 + 64
    dec = RC.RangeDecoder(path)
    assert np.array_equal(dec.decode(sym.size, cum, dtype=np.uint8), sym)


def test_streams_batch_equals_the_file_coder(tmp_path):
    """encode_streams (thread pool, in memory) writes exactly the bytes RangeEncoder(path).encode(...); close() writes;
    decode_streams reads them back; every thread count gives the same bytes."""
    rs = np.random.RandomState(12)
    for k, total in ((2, 4096), (2, 3000), (7, 1000), (256, 65536)):
        p = rs.dirichlet([0.4] * k)
        cum = RC.prob_to_cum_freq(p * 0.99 + 0.01 / k, resolution=total)
        streams = [rs.choice(k, size=n, p=p).astype(np.uint8) for n in (0, 1, 17, 4096, 30001)]
        blobs = RC.encode_streams(streams, cum, threads=3)
        assert blobs == RC.encode_streams(streams, cum, threads=1)
        for i, (s, b) in enumerate(zip(streams, blobs)):
            path = tmp_path / f"s{k}_{i}.bin"
            e = RC.RangeEncoder(str(path))
            e.encode(s, cum)
            e.close()
            assert path.read_bytes() == b and e.bytes_written == len(b)
            assert len(b) <= RC.load().tic_rc_max_encoded_bytes(len(s))
        back = RC.decode_streams(blobs, [len(s) for s in streams], cum, threads=2)
        assert all(np.array_equal(a, b) for a, b in zip(back, streams))
    # a 2-D array is one stream per row
    sym = (rs.rand(5, 1000) < 0.3).astype(np.uint8)
    assert RC.encode_streams(sym, [0, 2800, 4096]) == RC.encode_streams(list(sym), [0, 2800, 4096])


def test_carry_propagation_and_stream_tail(tmp_path):
    """Long runs of a near-certain symbol park 0xFF bytes behind the cache byte; a rare symbol then carries through
    them.  Round trips at extreme skews and totals, empty streams store nothing, trailing zero bytes are not stored."""
    rs = np.random.RandomState(3)
    for cum in ([0, 65535, 65536], [0, 1, 65536], [0, 4095, 4096], [0, 1, 2], [0, 255, 256, 65536]):
        k = len(cum) - 1
        w = np.diff(cum) / cum[-1]
        for n in (1, 2, 3, 1000, 50000):
            s = rs.choice(k, size=n, p=w).astype(np.uint8)
            if n >= 1000:
                s[rs.randint(0, n, size=5)] = int(np.argmin(w))  # force the rare symbol in
            b, = RC.encode_streams([s], cum)
            assert np.array_equal(RC.decode_streams([b], [n], cum)[0], s), (cum, n)
            assert not b or b[-1] != 0  # trailing zeros are implied, never stored
    assert RC.encode_streams([np.zeros(0, np.uint8)], [0, 1, 2]) == [b""]
    # the all-zeros code value: a stream of the first symbol under a dyadic table stores nothing at all
    assert RC.encode_streams([np.zeros(64, np.uint8)], [0, 2, 4]) == [b""]
    assert np.array_equal(RC.decode_streams([b""], [64], [0, 2, 4])[0], np.zeros(64, np.uint8))
    # totals above 2^16 are rejected (include/tic_rc_core.h), like any invalid table
    with pytest.raises(ValueError):
        RC.encode_streams([np.zeros(4, np.uint8)], [0, 1, 65537])
    e = RC.RangeEncoder(str(tmp_path / "t.bin"))
    with pytest.raises(ValueError):
        e.encode([0, 1], [0, 40000, 80000])
    e.close()


def test_segmented_calls(tmp_path):
    """include/tic_rc_core.h "SEGMENTED CALLS": a call of more than 32768 symbols on a fresh stream is a container of
    little-endian uint32 byte counts followed by independently coded 32768-symbol streams; shorter calls, and calls on a
    stream that already holds a plain call, are plain streams.  The decoder applies the same rule."""
    SEG = 32768
    rs = np.random.RandomState(21)
    lib = RC.load()
    for cum in ([0, 2252, 4096], [0, 700, 1000], [0, 10, 300, 301, 4096]):
        k = len(cum) - 1
        w = np.diff(cum) / cum[-1]
        for n in (SEG, SEG + 1, 2 * SEG, 3 * SEG + 17):
            s = rs.choice(k, size=n, p=w).astype(np.uint8)
            blob, = RC.encode_streams([s], cum, threads=2)
            assert len(blob) <= lib.tic_rc_max_encoded_bytes(n)
            nseg = -(-n // SEG) if n > SEG else 0
            if nseg == 0:
                plain = blob
            else:
                lens = np.frombuffer(blob[:4 * nseg], "<u4")
                assert 4 * nseg + int(lens.sum()) == len(blob)
                at = 4 * nseg
                for j, ln in enumerate(lens):
                    piece = s[j * SEG:(j + 1) * SEG]
                    assert blob[at:at + ln] == RC.encode_streams([piece], cum)[0]  # a plain stream of its own
                    at += int(ln)
            assert np.array_equal(RC.decode_streams([blob], [n], cum, threads=3)[0], s)
            path = tmp_path / "seg.bin"
            e = RC.RangeEncoder(str(path))
            e.encode(s, cum)
            e.close()
            assert path.read_bytes() == blob
    # several calls on one stream: segmented calls leave it fresh, the first plain call ends that
    cum = [0, 2252, 4096]
    calls = [(rs.rand(n) < 0.45).astype(np.uint8) for n in (2 * SEG + 5, SEG + 1, 100, SEG + 9, 7)]
    path = tmp_path / "mixed.bin"
    e = RC.RangeEncoder(str(path))
    for c in calls:
        e.encode(c, cum)
    e.close()
    raw = path.read_bytes()
    first = RC.encode_streams([calls[0]], cum)[0] + RC.encode_streams([calls[1]], cum)[0]
    assert raw[:len(first)] == first
    # ... and what follows is ONE plain stream of the remaining three calls
    rest = path.with_suffix(".rest")
    e = RC.RangeEncoder(str(rest))
    e.encode(np.concatenate(calls[2:])[:SEG], cum)       # (kept below the segment size: a plain call)
    e.encode(np.concatenate(calls[2:])[SEG:], cum)
    e.close()
    assert raw[len(first):] == rest.read_bytes()
    d = RC.RangeDecoder(str(path))
    for c in calls:
        assert np.array_equal(d.decode(len(c), cum, dtype=np.uint8), c)
    d.close()
    # a corrupt header cannot drive the decoder out of bounds: garbage decodes to garbage of the right length
    bad = b"\xff" * 64
    assert len(RC.decode_streams([bad], [SEG + 1], cum)[0]) == SEG + 1


def test_rangecoder_library_exports_every_declared_symbol():
    import ctypes
    import re
    from tf_image_compression_b200 import _lib as L
    header = (ROOT / "include" / "tic_rangecoder.h").read_text()
    declared = set(re.findall(r"\b(tic_rc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(RC._SIG), declared ^ set(RC._SIG)
    lib = ctypes.CDLL(str(L.RC_LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name


def test_prob_to_cum_freq_invariants():
    rs = np.random.RandomState(190)
    p0 = rs.dirichlet([0.1] * 50)
    c0 = RC.prob_to_cum_freq(p0, 1024)
    p1 = RC.cum_freq_to_prob(c0)
    assert c0[0] == 0 and c0[-1] == 1024 and len(c0) == 51
    assert np.all(np.diff(c0)[p0 > 0] > 0)
    assert np.isclose(p1.sum(), 1.0)
    assert RC.prob_to_cum_freq(p1, 1024) == c0
    assert RC.prob_to_cum_freq([0.5, 0.25, 0.25], resolution=8) == [0, 4, 6, 8]
    z = RC.prob_to_cum_freq([0.5, 0.0, 0.25, 0.25, 0.0, 0.0], resolution=8)
    assert [z[0]] + [z[i + 1] for i, p in enumerate([0.5, 0.0, 0.25, 0.25, 0.0, 0.0]) if p > 0] == [0, 4, 6, 8]


@pytest.mark.skipif(not REF_TEST.exists(), reason="/root/reference is not mounted (GPU box)")
def test_reference_suite_passes_unmodified(tmp_path):
    """Run the reference's own other/test_range_coder.py against this module, aliased as `range_coder`."""
    shim = tmp_path / "range_coder"
    shim.mkdir()
    (shim / "__init__.py").write_text(
        "import sys\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from tf_image_compression_b200.range_coder import RangeEncoder, RangeDecoder, prob_to_cum_freq, cum_freq_to_prob\n")
    env = dict(os.environ, PYTHONPATH=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", str(REF_TEST)], env=env,
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
