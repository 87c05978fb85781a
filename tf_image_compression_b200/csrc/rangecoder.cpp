// Host entropy coder behind include/tic_rangecoder.h.  The arithmetic lives in include/tic_rc_core.h (single source
// with the CUDA entropy stage): a carry-propagating byte-wise range coder, static cumulative-frequency tables supplied
// per call.  This file adds files, buffers, table checks, whole-stream batches on a thread pool, prob_to_cum_freq and
// the CRC-32C of the checkpoint reader (host-only, so reading a checkpoint never needs the CUDA library).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/tic_rangecoder.h"
#include "../../include/tic_rc_core.h"

namespace {

int check_table(const uint32_t* cum, int n_cum) {
  if (!cum || n_cum < 2) return TIC_RC_ERR_TABLE;  // [] and [0] are invalid tables
  if (cum[0] != 0) return TIC_RC_ERR_TABLE;
  for (int i = 1; i < n_cum; ++i)
    if (cum[i] < cum[i - 1]) return TIC_RC_ERR_TABLE;
  if (cum[n_cum - 1] == 0 || cum[n_cum - 1] > TIC_RC_MAX_TOTAL) return TIC_RC_ERR_TABLE;
  return TIC_RC_OK;
}

inline bool is_pow2(uint32_t v) { return (v & (v - 1)) == 0; }
inline int log2u(uint32_t v) {
  int k = 0;
  while ((1u << k) < v) ++k;
  return k;
}

// Byte sink that does not store the trailing zero bytes of the stream: zeros are counted and only written once a
// non-zero byte follows them.
struct FileSink {
  FILE* f = nullptr;
  std::vector<uint8_t> buf;
  uint64_t zrun = 0;
  int64_t stored = 0;
  bool io_error = false;
  void raw(uint8_t b) {
    buf.push_back(b);
    ++stored;
    if (buf.size() >= (1u << 16)) flush();
  }
  void put(uint8_t b) {
    if (b == 0) {
      ++zrun;
      return;
    }
    for (; zrun > 0; --zrun) raw(0);
    raw(b);
  }
  void flush() {
    if (!buf.empty() && f) {
      if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) io_error = true;
    }
    buf.clear();
  }
  void put_raw(const uint8_t* src, size_t cnt) {
    for (; zrun > 0; --zrun) raw(0);  // (a container only ever follows a fresh state: nothing is pending)
    for (size_t i = 0; i < cnt; ++i) raw(src[i]);
  }
};

// Sink into caller memory; `len` ends after the last non-zero byte (bytes are written as they come, so zeros inside
// the stream are in place and the trailing ones are simply not counted).
struct MemSink {
  uint8_t* p;
  int64_t cap, pos = 0, len = 0;
  bool overflow = false;
  MemSink(uint8_t* p_, int64_t cap_) : p(p_), cap(cap_) {}
  void put(uint8_t b) {
    if (pos >= cap) {
      overflow = true;
      return;
    }
    p[pos++] = b;
    if (b) len = pos;
  }
  // container bytes: stored as they are (a container is never zero-stripped; what follows it starts a fresh stream)
  void put_raw(const uint8_t* src, size_t cnt) {
    if (pos + (int64_t)cnt > cap) {
      overflow = true;
      return;
    }
    memcpy(p + pos, src, cnt);
    pos += (int64_t)cnt;
    len = pos;
  }
};

struct FileSource {
  FILE* f = nullptr;
  std::vector<uint8_t> buf;
  size_t pos = 0;
  uint32_t get() {
    if (pos >= buf.size()) {
      buf.resize(1 << 16);
      const size_t got = f ? fread(buf.data(), 1, buf.size(), f) : 0;
      buf.resize(got);
      pos = 0;
      if (got == 0) return 0;  // zero extension past EOF
    }
    return buf[pos++];
  }
  // n raw bytes (container header / segment bodies); short reads are zero-filled like reads past EOF
  void read_raw(uint8_t* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) dst[i] = (uint8_t)get();
  }
};

struct MemSource {
  const uint8_t* p;
  int64_t n, pos = 0;
  MemSource(const uint8_t* p_, int64_t n_) : p(p_), n(n_) {}
  uint32_t get() { return pos < n ? p[pos++] : 0u; }
  void read_raw(uint8_t* dst, size_t cnt) {
    for (size_t i = 0; i < cnt; ++i) dst[i] = (uint8_t)get();
  }
};

template <typename T, class Sink>
int encode_symbols(tic_rc_enc_state* st, Sink& out, const T* sym, int64_t n, const uint32_t* cum, int n_cum) {
  const uint32_t total = cum[n_cum - 1];
  const int64_t nsym = n_cum - 1;
  const bool p2 = is_pow2(total);
  const int k = p2 ? log2u(total) : 0;
  if (nsym == 2 && p2 && cum[1] != 0 && cum[1] != total) {
    // binary alphabet with a power-of-two total (quan_scale = 2, resolution = 4096): the path the GPU stage runs
    const uint32_t c1 = cum[1];
    for (int64_t i = 0; i < n; ++i) {
      const int64_t s = (int64_t)sym[i];
      if (s < 0 || s > 1) return TIC_RC_ERR_SYMBOL;
      tic_rc_enc_bit(st, out, (uint32_t)s, k, c1);
    }
    return TIC_RC_OK;
  }
  for (int64_t i = 0; i < n; ++i) {
    const int64_t s = (int64_t)sym[i];
    if (s < 0 || s >= nsym) return TIC_RC_ERR_SYMBOL;
    const uint32_t lo = cum[s], hi = cum[s + 1];
    if (hi == lo) return TIC_RC_ERR_SYMBOL;  // symbols with zero probability cannot be encoded
    const uint32_t r = p2 ? st->range >> k : st->range / total;
    tic_rc_enc_step(st, out, r, lo, hi);
  }
  return TIC_RC_OK;
}

template <typename T, class Source>
int decode_symbols(tic_rc_dec_state* st, Source& in, T* out, int64_t n, const uint32_t* cum, int n_cum) {
  tic_rc_dec_prime(st, in);
  const uint32_t total = cum[n_cum - 1];
  const bool p2 = is_pow2(total);
  const int k = p2 ? log2u(total) : 0;
  if (n_cum == 3) {
    // binary alphabet (every shipped config: quan_scale = 2): one compare instead of the division and the search
    const uint32_t c1 = cum[1];
    if (p2 && c1 != 0 && c1 != total) {
      // the shipped case, branch-free on the (near coin-flip) symbol value; same arithmetic as the loop below
      uint32_t range = st->range, code = st->code;
      for (int64_t i = 0; i < n; ++i) {
        const uint32_t r = range >> k;
        const uint32_t t = r * c1;
        const uint32_t m = 0u - (uint32_t)(code >= t);
        out[i] = (T)(m & 1u);
        code -= t & m;
        range = (t & ~m) | (((r << k) - t) & m);
        while (range < TIC_RC_TOP) {
          code = (code << 8) | in.get();
          range <<= 8;
        }
      }
      st->range = range;
      st->code = code;
      return TIC_RC_OK;
    }
    for (int64_t i = 0; i < n; ++i) {
      const uint32_t r = p2 ? st->range >> k : st->range / total;
      const uint32_t t = r * c1;
      // zero-width symbols: c1 == 0 -> always 1; c1 == total -> 1 only for corrupt streams (clamped to 0)
      const bool one = c1 == total ? false : st->code >= t;
      out[i] = (T)(one ? 1 : 0);
      tic_rc_dec_step(st, in, r, one ? c1 : 0u, one ? total : c1);
    }
    return TIC_RC_OK;
  }
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t r = p2 ? st->range >> k : st->range / total;
    const uint32_t v = tic_rc_dec_target(st, r, total);
    // last entry with cum[s] <= v (skips zero-width symbols)
    const uint32_t* it = std::upper_bound(cum, cum + n_cum, v);
    const int64_t s = (it - cum) - 1;
    out[i] = (T)s;
    tic_rc_dec_step(st, in, r, cum[s], cum[s + 1]);
  }
  return TIC_RC_OK;
}

template <class Fn>
void parallel_for(int64_t n, int n_threads, Fn fn) {
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
  if (n_threads == 1) {
    for (int64_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> pool;
  pool.reserve(n_threads);
  for (int t = 0; t < n_threads; ++t)
    pool.emplace_back([&] {
      for (;;) {
        const int64_t i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto& th : pool) th.join();
}

// ---- segmented calls (include/tic_rc_core.h): a call of more than TIC_RC_SEGMENT_SYMBOLS symbols on a fresh stream ----
template <typename T>
int encode_segment(const T* sym, int64_t n, const uint32_t* cum, int n_cum, std::vector<uint8_t>* out) {
  // worst-case scratch once per worker thread (a fresh zero-filled 66 KB vector per segment cost more than the coding)
  thread_local std::unique_ptr<uint8_t[]> scratch;
  thread_local int64_t scratch_cap = 0;
  const int64_t need = tic_rc_plain_bound(n);
  if (scratch_cap < need) {
    scratch.reset(new uint8_t[(size_t)need]);
    scratch_cap = need;
  }
  tic_rc_enc_state st;
  tic_rc_enc_init(&st);
  MemSink sink(scratch.get(), need);
  int rc = encode_symbols(&st, sink, sym, n, cum, n_cum);
  if (rc != TIC_RC_OK) return rc;
  tic_rc_enc_finish(&st, sink);
  if (sink.overflow) return TIC_RC_ERR_IO;
  out->assign(scratch.get(), scratch.get() + sink.len);
  return TIC_RC_OK;
}

// header (nseg little-endian uint32 byte counts) + bodies into any sink with put_raw
template <typename T, class Sink>
int encode_segmented(Sink& out, const T* sym, int64_t n, const uint32_t* cum, int n_cum, int n_threads) {
  const int64_t nseg = tic_rc_segments(n);
  std::vector<std::vector<uint8_t>> body((size_t)nseg);
  std::atomic<int> status{TIC_RC_OK};
  parallel_for(nseg, n_threads, [&](int64_t j) {
    const int64_t s0 = j * TIC_RC_SEGMENT_SYMBOLS, len = std::min<int64_t>(TIC_RC_SEGMENT_SYMBOLS, n - s0);
    const int r = encode_segment(sym + s0, len, cum, n_cum, &body[(size_t)j]);
    if (r != TIC_RC_OK) status.store(r);
  });
  if (status.load() != TIC_RC_OK) return status.load();
  std::vector<uint8_t> hdr((size_t)nseg * 4);
  for (int64_t j = 0; j < nseg; ++j) {
    const uint32_t v = (uint32_t)body[(size_t)j].size();
    hdr[4 * j] = (uint8_t)v;
    hdr[4 * j + 1] = (uint8_t)(v >> 8);
    hdr[4 * j + 2] = (uint8_t)(v >> 16);
    hdr[4 * j + 3] = (uint8_t)(v >> 24);
  }
  out.put_raw(hdr.data(), hdr.size());
  for (auto& b : body) out.put_raw(b.data(), b.size());
  return TIC_RC_OK;
}

template <typename T, class Source>
int decode_segmented(Source& in, T* out, int64_t n, const uint32_t* cum, int n_cum, int n_threads) {
  const int64_t nseg = tic_rc_segments(n);
  std::vector<uint8_t> hdr((size_t)nseg * 4);
  in.read_raw(hdr.data(), hdr.size());
  std::vector<std::vector<uint8_t>> body((size_t)nseg);
  for (int64_t j = 0; j < nseg; ++j) {
    uint32_t v = (uint32_t)hdr[4 * j] | ((uint32_t)hdr[4 * j + 1] << 8) | ((uint32_t)hdr[4 * j + 2] << 16) | ((uint32_t)hdr[4 * j + 3] << 24);
    v = (uint32_t)std::min<int64_t>(v, tic_rc_plain_bound(TIC_RC_SEGMENT_SYMBOLS));  // corrupt header: stay bounded
    body[(size_t)j].resize(v);
    in.read_raw(body[(size_t)j].data(), v);
  }
  parallel_for(nseg, n_threads, [&](int64_t j) {
    const int64_t s0 = j * TIC_RC_SEGMENT_SYMBOLS, len = std::min<int64_t>(TIC_RC_SEGMENT_SYMBOLS, n - s0);
    tic_rc_dec_state st;
    tic_rc_dec_init(&st);
    MemSource src(body[(size_t)j].data(), (int64_t)body[(size_t)j].size());
    decode_symbols(&st, src, out + s0, len, cum, n_cum);
  });
  return TIC_RC_OK;
}

// One encode() call: segmented when the stream is fresh and the call is long, plain otherwise.
template <typename T, class Sink>
int encode_call(tic_rc_enc_state* st, bool* fresh, Sink& out, const T* sym, int64_t n, const uint32_t* cum, int n_cum, int n_threads) {
  if (*fresh && tic_rc_segments(n) > 0) {
    // validate the symbols before anything is written (a failing call must not leave half a container behind)
    const int64_t nsym = n_cum - 1;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t s = (int64_t)sym[i];
      if (s < 0 || s >= nsym || cum[s + 1] == cum[s]) return TIC_RC_ERR_SYMBOL;
    }
    return encode_segmented(out, sym, n, cum, n_cum, n_threads);  // the stream stays fresh
  }
  if (n > 0) *fresh = false;
  return encode_symbols(st, out, sym, n, cum, n_cum);
}

template <typename T, class Source>
int decode_call(tic_rc_dec_state* st, bool* fresh, Source& in, T* out, int64_t n, const uint32_t* cum, int n_cum, int n_threads) {
  if (*fresh && tic_rc_segments(n) > 0) return decode_segmented(in, out, n, cum, n_cum, n_threads);
  if (n > 0) *fresh = false;
  return decode_symbols(st, in, out, n, cum, n_cum);
}

}  // namespace

struct tic_rc_encoder {
  tic_rc_enc_state st;
  FileSink out;
  bool fresh = true;
};

struct tic_rc_decoder {
  tic_rc_dec_state st;
  FileSource in;
  bool fresh = true;
};

extern "C" {

int tic_rc_encoder_open(tic_rc_encoder** out, const char* path) {
  if (!out || !path) return TIC_RC_ERR_IO;
  *out = nullptr;
  FILE* f = fopen(path, "wb");
  if (!f) return TIC_RC_ERR_IO;
  tic_rc_encoder* e = new tic_rc_encoder();
  tic_rc_enc_init(&e->st);
  e->out.f = f;
  *out = e;
  return TIC_RC_OK;
}

int tic_rc_encode_u8(tic_rc_encoder* e, const uint8_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!e || !e->out.f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return encode_call(&e->st, &e->fresh, e->out, symbols, n, cum_freq, n_cum, 0);
}

int tic_rc_encode_i32(tic_rc_encoder* e, const int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!e || !e->out.f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return encode_call(&e->st, &e->fresh, e->out, symbols, n, cum_freq, n_cum, 0);
}

int tic_rc_encoder_close(tic_rc_encoder* e) {
  if (!e) return TIC_RC_ERR_CLOSED;
  if (!e->out.f) return TIC_RC_OK;
  tic_rc_enc_finish(&e->st, e->out);
  e->out.flush();  // pending zeros are the stream's trailing zeros: not stored
  const bool bad = e->out.io_error || fclose(e->out.f) != 0;
  e->out.f = nullptr;
  return bad ? TIC_RC_ERR_IO : TIC_RC_OK;
}

void tic_rc_encoder_free(tic_rc_encoder* e) {
  if (!e) return;
  tic_rc_encoder_close(e);
  delete e;
}

int64_t tic_rc_encoder_bytes(const tic_rc_encoder* e) { return e ? e->out.stored : 0; }

int tic_rc_decoder_open(tic_rc_decoder** out, const char* path) {
  if (!out || !path) return TIC_RC_ERR_IO;
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return TIC_RC_ERR_IO;
  tic_rc_decoder* d = new tic_rc_decoder();
  tic_rc_dec_init(&d->st);
  d->in.f = f;
  *out = d;
  return TIC_RC_OK;
}

int tic_rc_decode_u8(tic_rc_decoder* d, uint8_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!d || !d->in.f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n_cum - 1 > 256) return TIC_RC_ERR_TABLE;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return decode_call(&d->st, &d->fresh, d->in, symbols, n, cum_freq, n_cum, 0);
}

int tic_rc_decode_i32(tic_rc_decoder* d, int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!d || !d->in.f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return decode_call(&d->st, &d->fresh, d->in, symbols, n, cum_freq, n_cum, 0);
}

int tic_rc_decoder_close(tic_rc_decoder* d) {
  if (!d) return TIC_RC_ERR_CLOSED;
  if (d->in.f) fclose(d->in.f);
  d->in.f = nullptr;
  return TIC_RC_OK;
}

void tic_rc_decoder_free(tic_rc_decoder* d) {
  if (!d) return;
  tic_rc_decoder_close(d);
  delete d;
}

int64_t tic_rc_max_encoded_bytes(int64_t n_symbols) { return tic_rc_bound(n_symbols < 0 ? 0 : n_symbols); }

int tic_rc_encode_streams(const uint8_t* symbols, const int64_t* sym_offsets, int64_t n_streams, const uint32_t* cum_freq,
                          int n_cum, uint8_t* out, const int64_t* out_offsets, int64_t* out_bytes, int n_threads) {
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n_streams < 0 || (n_streams > 0 && (!symbols || !sym_offsets || !out || !out_offsets || !out_bytes))) return TIC_RC_ERR_SYMBOL;
  // work items: every segment of every long stream, every short stream as a whole — all on one thread pool
  struct Item {
    int64_t stream, seg, s0, len;
  };
  std::vector<Item> items;
  std::vector<int64_t> first(n_streams + 1, 0);
  for (int64_t i = 0; i < n_streams; ++i) {
    const int64_t n = sym_offsets[i + 1] - sym_offsets[i], nseg = tic_rc_segments(n);
    first[i] = (int64_t)items.size();
    if (nseg == 0) {
      items.push_back({i, -1, 0, n});
    } else {
      for (int64_t j = 0; j < nseg; ++j)
        items.push_back({i, j, j * TIC_RC_SEGMENT_SYMBOLS, std::min<int64_t>(TIC_RC_SEGMENT_SYMBOLS, n - j * TIC_RC_SEGMENT_SYMBOLS)});
    }
  }
  first[n_streams] = (int64_t)items.size();
  std::vector<std::vector<uint8_t>> body(items.size());
  std::atomic<int> status{TIC_RC_OK};
  parallel_for((int64_t)items.size(), n_threads, [&](int64_t k) {
    const Item& it = items[(size_t)k];
    const int r = encode_segment(symbols + sym_offsets[it.stream] + it.s0, it.len, cum_freq, n_cum, &body[(size_t)k]);
    if (r != TIC_RC_OK) status.store(r);
  });
  if (status.load() != TIC_RC_OK) {
    for (int64_t i = 0; i < n_streams; ++i) out_bytes[i] = 0;
    return status.load();
  }
  parallel_for(n_streams, n_threads, [&](int64_t i) {
    MemSink sink(out + out_offsets[i], out_offsets[i + 1] - out_offsets[i]);
    const int64_t k0 = first[i], k1 = first[i + 1];
    if (items[(size_t)k0].seg >= 0) {
      for (int64_t k = k0; k < k1; ++k) {
        const uint32_t v = (uint32_t)body[(size_t)k].size();
        const uint8_t h4[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
        sink.put_raw(h4, 4);
      }
    }
    for (int64_t k = k0; k < k1; ++k) sink.put_raw(body[(size_t)k].data(), body[(size_t)k].size());
    out_bytes[i] = sink.overflow ? 0 : sink.pos;
    if (sink.overflow) status.store(TIC_RC_ERR_IO);
  });
  return status.load();
}

int tic_rc_decode_streams(const uint8_t* in, const int64_t* in_offsets, const int64_t* in_bytes, int64_t n_streams,
                          const uint32_t* cum_freq, int n_cum, uint8_t* symbols, const int64_t* sym_offsets, int n_threads) {
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n_cum - 1 > 256) return TIC_RC_ERR_TABLE;
  if (n_streams < 0 || (n_streams > 0 && (!in || !in_offsets || !in_bytes || !symbols || !sym_offsets))) return TIC_RC_ERR_SYMBOL;
  struct Item {
    const uint8_t* p;
    int64_t bytes;
    uint8_t* out;
    int64_t len;
  };
  std::vector<Item> items;
  for (int64_t i = 0; i < n_streams; ++i) {
    const int64_t n = sym_offsets[i + 1] - sym_offsets[i], nseg = tic_rc_segments(n);
    const uint8_t* base = in + in_offsets[i];
    const int64_t have = in_bytes[i];
    if (nseg == 0) {
      items.push_back({base, have, symbols + sym_offsets[i], n});
      continue;
    }
    int64_t pos = 4 * nseg;
    for (int64_t j = 0; j < nseg; ++j) {
      uint32_t v = 0;
      for (int b = 0; b < 4; ++b)
        if (4 * j + b < have) v |= (uint32_t)base[4 * j + b] << (8 * b);
      const int64_t avail = std::max<int64_t>(0, std::min<int64_t>(v, have - pos));  // corrupt header: stay inside the stream
      items.push_back({base + std::min(pos, have), avail, symbols + sym_offsets[i] + j * TIC_RC_SEGMENT_SYMBOLS,
                       std::min<int64_t>(TIC_RC_SEGMENT_SYMBOLS, n - j * TIC_RC_SEGMENT_SYMBOLS)});
      pos += v;
    }
  }
  parallel_for((int64_t)items.size(), n_threads, [&](int64_t k) {
    const Item& it = items[(size_t)k];
    tic_rc_dec_state st;
    tic_rc_dec_init(&st);
    MemSource src(it.p, it.bytes);
    decode_symbols(&st, src, it.out, it.len, cum_freq, n_cum);
  });
  return TIC_RC_OK;
}

int tic_rc_prob_to_cum_freq(const double* prob, int n, uint32_t resolution, uint32_t* cum_freq) {
  if (!prob || !cum_freq || n <= 0 || resolution == 0) return TIC_RC_ERR_TABLE;
  double sum = 0.0;
  int nz = 0;
  for (int i = 0; i < n; ++i) {
    if (!(prob[i] >= 0.0)) return TIC_RC_ERR_TABLE;
    sum += prob[i];
    nz += prob[i] > 0.0;
  }
  if (!(sum > 0.0) || (uint32_t)nz > resolution) return TIC_RC_ERR_TABLE;
  // floor of the scaled probabilities, at least 1 for every non-zero entry; the remainder goes to the
  // largest fractional parts (ties: lowest index), any excess comes off the largest counts
  std::vector<int64_t> freq(n, 0);
  std::vector<double> frac(n, 0.0);
  int64_t used = 0;
  for (int i = 0; i < n; ++i) {
    if (prob[i] > 0.0) {
      const double x = prob[i] / sum * (double)resolution;
      int64_t fl = (int64_t)x;
      frac[i] = x - (double)fl;
      if (fl < 1) {
        fl = 1;
        frac[i] = 0.0;
      }
      freq[i] = fl;
      used += fl;
    }
  }
  std::vector<int> order;
  for (int i = 0; i < n; ++i)
    if (prob[i] > 0.0) order.push_back(i);
  if (used < (int64_t)resolution) {
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return frac[a] > frac[b]; });
    size_t k = 0;
    while (used < (int64_t)resolution) {
      ++freq[order[k % order.size()]];
      ++used;
      ++k;
    }
  } else if (used > (int64_t)resolution) {
    while (used > (int64_t)resolution) {
      int best = -1;
      for (int i : order)
        if (freq[i] > 1 && (best < 0 || freq[i] > freq[best])) best = i;
      if (best < 0) return TIC_RC_ERR_TABLE;
      --freq[best];
      --used;
    }
  }
  cum_freq[0] = 0;
  for (int i = 0; i < n; ++i) cum_freq[i + 1] = cum_freq[i] + (uint32_t)freq[i];
  return TIC_RC_OK;
}

uint32_t tic_rc_crc32c(const void* data, uint64_t n) {
  static uint32_t table[8][256];
  static const bool ready = [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xffu];
    return true;
  }();
  (void)ready;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = 0xffffffffu;
  while (n >= 8) {  // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xffu] ^ table[6][(lo >> 8) & 0xffu] ^ table[5][(lo >> 16) & 0xffu] ^ table[4][lo >> 24] ^
        table[3][hi & 0xffu] ^ table[2][(hi >> 8) & 0xffu] ^ table[1][(hi >> 16) & 0xffu] ^ table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ table[0][(c ^ *p++) & 0xffu];
  return c ^ 0xffffffffu;
}

}  // extern "C"
