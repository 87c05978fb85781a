// Host entropy coder behind include/tic_rangecoder.h.  The arithmetic lives in include/tic_rc_core.h (single source
// with the CUDA entropy stage): a carry-propagating byte-wise range coder, static cumulative-frequency tables supplied
// per call.  This file adds files, buffers, table checks, whole-stream batches on a thread pool, prob_to_cum_freq and
// the CRC-32C of the checkpoint reader (host-only, so reading a checkpoint never needs the CUDA library).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
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
};

struct MemSource {
  const uint8_t* p;
  int64_t n, pos = 0;
  MemSource(const uint8_t* p_, int64_t n_) : p(p_), n(n_) {}
  uint32_t get() { return pos < n ? p[pos++] : 0u; }
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

}  // namespace

struct tic_rc_encoder {
  tic_rc_enc_state st;
  FileSink out;
};

struct tic_rc_decoder {
  tic_rc_dec_state st;
  FileSource in;
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
  return encode_symbols(&e->st, e->out, symbols, n, cum_freq, n_cum);
}

int tic_rc_encode_i32(tic_rc_encoder* e, const int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!e || !e->out.f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return encode_symbols(&e->st, e->out, symbols, n, cum_freq, n_cum);
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
  return decode_symbols(&d->st, d->in, symbols, n, cum_freq, n_cum);
}

int tic_rc_decode_i32(tic_rc_decoder* d, int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!d || !d->in.f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return decode_symbols(&d->st, d->in, symbols, n, cum_freq, n_cum);
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
  std::atomic<int> status{TIC_RC_OK};
  parallel_for(n_streams, n_threads, [&](int64_t i) {
    tic_rc_enc_state st;
    tic_rc_enc_init(&st);
    MemSink sink(out + out_offsets[i], out_offsets[i + 1] - out_offsets[i]);
    int r = encode_symbols(&st, sink, symbols + sym_offsets[i], sym_offsets[i + 1] - sym_offsets[i], cum_freq, n_cum);
    if (r == TIC_RC_OK) {
      tic_rc_enc_finish(&st, sink);
      if (sink.overflow) r = TIC_RC_ERR_IO;
    }
    out_bytes[i] = r == TIC_RC_OK ? sink.len : 0;
    if (r != TIC_RC_OK) status.store(r);
  });
  return status.load();
}

int tic_rc_decode_streams(const uint8_t* in, const int64_t* in_offsets, const int64_t* in_bytes, int64_t n_streams,
                          const uint32_t* cum_freq, int n_cum, uint8_t* symbols, const int64_t* sym_offsets, int n_threads) {
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n_cum - 1 > 256) return TIC_RC_ERR_TABLE;
  if (n_streams < 0 || (n_streams > 0 && (!in || !in_offsets || !in_bytes || !symbols || !sym_offsets))) return TIC_RC_ERR_SYMBOL;
  parallel_for(n_streams, n_threads, [&](int64_t i) {
    tic_rc_dec_state st;
    tic_rc_dec_init(&st);
    MemSource src(in + in_offsets[i], in_bytes[i]);
    decode_symbols(&st, src, symbols + sym_offsets[i], sym_offsets[i + 1] - sym_offsets[i], cum_freq, n_cum);
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
