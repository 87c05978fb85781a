// Host entropy coder behind include/tic_rangecoder.h: integer arithmetic coding with 32-bit code values
// (the published "low / high / underflow bits" scheme), static cumulative-frequency tables supplied per
// call, MSB-first bit stream.  The decoder zero-extends the stream past EOF, which lets close() write the
// shortest tail whose zero extension lies inside the final interval (for a dyadic source that ends on a
// byte boundary: nothing, so the reference's known-answer vector is exactly one byte per 8 bits).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/tic_rangecoder.h"

namespace {

constexpr uint64_t kMax = 0xFFFFFFFFull;       // largest code value
constexpr uint64_t kQuarter = 0x40000000ull;
constexpr uint64_t kHalf = 0x80000000ull;
constexpr uint64_t kThreeQuarters = 0xC0000000ull;
constexpr uint64_t kMaxTotal = kQuarter;       // frequency totals above this lose the decodability guarantee

int check_table(const uint32_t* cum, int n_cum) {
  if (!cum || n_cum < 2) return TIC_RC_ERR_TABLE;  // [] and [0] are invalid tables
  if (cum[0] != 0) return TIC_RC_ERR_TABLE;
  for (int i = 1; i < n_cum; ++i)
    if (cum[i] < cum[i - 1]) return TIC_RC_ERR_TABLE;
  if (cum[n_cum - 1] == 0 || (uint64_t)cum[n_cum - 1] > kMaxTotal) return TIC_RC_ERR_TABLE;
  return TIC_RC_OK;
}

}  // namespace

struct tic_rc_encoder {
  FILE* f = nullptr;
  uint64_t low = 0, high = kMax;
  uint64_t pending = 0;
  uint8_t cur = 0;
  int nbits = 0;
  int64_t bytes = 0;
  std::vector<uint8_t> buf;
  bool io_error = false;

  void put_byte(uint8_t b) {
    buf.push_back(b);
    ++bytes;
    if (buf.size() >= (1u << 16)) flush_buf();
  }
  void flush_buf() {
    if (!buf.empty() && f) {
      if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) io_error = true;
      buf.clear();
    }
  }
  void put_bit(int b) {
    cur = (uint8_t)((cur << 1) | (b & 1));
    if (++nbits == 8) {
      put_byte(cur);
      cur = 0;
      nbits = 0;
    }
  }
  void put_bit_plus_pending(int b) {
    put_bit(b);
    for (; pending > 0; --pending) put_bit(!b);
  }
  template <typename T>
  int encode(const T* sym, int64_t n, const uint32_t* cum, int n_cum) {
    const uint64_t total = cum[n_cum - 1];
    const int64_t nsym = n_cum - 1;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t s = (int64_t)sym[i];
      if (s < 0 || s >= nsym) return TIC_RC_ERR_SYMBOL;
      const uint64_t lo = cum[s], hi = cum[s + 1];
      if (hi == lo) return TIC_RC_ERR_SYMBOL;  // symbols with zero probability cannot be encoded
      const uint64_t range = high - low + 1;
      high = low + range * hi / total - 1;
      low = low + range * lo / total;
      for (;;) {
        if (high < kHalf) {
          put_bit_plus_pending(0);
        } else if (low >= kHalf) {
          put_bit_plus_pending(1);
        } else if (low >= kQuarter && high < kThreeQuarters) {
          ++pending;
          low -= kQuarter;
          high -= kQuarter;
        } else {
          break;
        }
        high = ((high << 1) | 1) & kMax;
        low = (low << 1) & kMax;
      }
    }
    return TIC_RC_OK;
  }
  void terminate() {
    if (pending > 0) {
      // interval straddles the middle: one deciding bit plus the parked underflow bits
      ++pending;
      put_bit_plus_pending(low < kQuarter ? 0 : 1);
    } else if (low != 0) {
      // shortest prefix whose zero extension lies in [low, high]
      for (int k = 1; k <= 32; ++k) {
        const uint64_t unit = 1ull << (32 - k);
        const uint64_t cand = (low + unit - 1) / unit * unit;
        if (cand <= high) {
          for (int b = 31; b >= 32 - k; --b) put_bit((int)((cand >> b) & 1));
          break;
        }
      }
    }
    if (nbits > 0) {
      put_byte((uint8_t)(cur << (8 - nbits)));
      cur = 0;
      nbits = 0;
    }
  }
};

struct tic_rc_decoder {
  FILE* f = nullptr;
  uint64_t low = 0, high = kMax, value = 0;
  uint8_t cur = 0;
  int nbits = 0;
  bool primed = false;
  std::vector<uint8_t> buf;
  size_t pos = 0;

  int get_bit() {
    if (nbits == 0) {
      if (pos >= buf.size()) {
        buf.resize(1 << 16);
        const size_t got = f ? fread(buf.data(), 1, buf.size(), f) : 0;
        buf.resize(got);
        pos = 0;
      }
      cur = pos < buf.size() ? buf[pos++] : 0;  // zero extension past EOF
      nbits = 8;
    }
    --nbits;
    return (cur >> nbits) & 1;
  }
  template <typename T>
  int decode(T* out, int64_t n, const uint32_t* cum, int n_cum) {
    if (!primed) {
      for (int i = 0; i < 32; ++i) value = (value << 1) | (uint64_t)get_bit();
      primed = true;
    }
    const uint64_t total = cum[n_cum - 1];
    for (int64_t i = 0; i < n; ++i) {
      const uint64_t range = high - low + 1;
      uint64_t scaled = ((value - low + 1) * total - 1) / range;
      if (scaled >= total) scaled = total - 1;  // corrupt stream: stay inside the table
      // last entry with cum[s] <= scaled (skips zero-width symbols)
      const uint32_t* it = std::upper_bound(cum, cum + n_cum, (uint32_t)scaled);
      const int64_t s = (it - cum) - 1;
      out[i] = (T)s;
      const uint64_t lo = cum[s], hi = cum[s + 1];
      high = low + range * hi / total - 1;
      low = low + range * lo / total;
      for (;;) {
        if (high < kHalf) {
        } else if (low >= kHalf) {
          value -= kHalf;
          low -= kHalf;
          high -= kHalf;
        } else if (low >= kQuarter && high < kThreeQuarters) {
          value -= kQuarter;
          low -= kQuarter;
          high -= kQuarter;
        } else {
          break;
        }
        low <<= 1;
        high = (high << 1) | 1;
        value = ((value << 1) | (uint64_t)get_bit()) & kMax;
      }
    }
    return TIC_RC_OK;
  }
};

extern "C" {

int tic_rc_encoder_open(tic_rc_encoder** out, const char* path) {
  if (!out || !path) return TIC_RC_ERR_IO;
  *out = nullptr;
  FILE* f = fopen(path, "wb");
  if (!f) return TIC_RC_ERR_IO;
  tic_rc_encoder* e = new tic_rc_encoder();
  e->f = f;
  *out = e;
  return TIC_RC_OK;
}

int tic_rc_encode_u8(tic_rc_encoder* e, const uint8_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!e || !e->f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return e->encode(symbols, n, cum_freq, n_cum);
}

int tic_rc_encode_i32(tic_rc_encoder* e, const int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!e || !e->f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return e->encode(symbols, n, cum_freq, n_cum);
}

int tic_rc_encoder_close(tic_rc_encoder* e) {
  if (!e) return TIC_RC_ERR_CLOSED;
  if (!e->f) return TIC_RC_OK;
  e->terminate();
  e->flush_buf();
  const bool bad = e->io_error || fclose(e->f) != 0;
  e->f = nullptr;
  return bad ? TIC_RC_ERR_IO : TIC_RC_OK;
}

void tic_rc_encoder_free(tic_rc_encoder* e) {
  if (!e) return;
  tic_rc_encoder_close(e);
  delete e;
}

int64_t tic_rc_encoder_bytes(const tic_rc_encoder* e) { return e ? e->bytes : 0; }

int tic_rc_decoder_open(tic_rc_decoder** out, const char* path) {
  if (!out || !path) return TIC_RC_ERR_IO;
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return TIC_RC_ERR_IO;
  tic_rc_decoder* d = new tic_rc_decoder();
  d->f = f;
  *out = d;
  return TIC_RC_OK;
}

int tic_rc_decode_u8(tic_rc_decoder* d, uint8_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!d || !d->f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n_cum - 1 > 256) return TIC_RC_ERR_TABLE;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return d->decode(symbols, n, cum_freq, n_cum);
}

int tic_rc_decode_i32(tic_rc_decoder* d, int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum) {
  if (!d || !d->f) return TIC_RC_ERR_CLOSED;
  int rc = check_table(cum_freq, n_cum);
  if (rc != TIC_RC_OK) return rc;
  if (n > 0 && !symbols) return TIC_RC_ERR_SYMBOL;
  return d->decode(symbols, n, cum_freq, n_cum);
}

int tic_rc_decoder_close(tic_rc_decoder* d) {
  if (!d) return TIC_RC_ERR_CLOSED;
  if (d->f) fclose(d->f);
  d->f = nullptr;
  return TIC_RC_OK;
}

void tic_rc_decoder_free(tic_rc_decoder* d) {
  if (!d) return;
  tic_rc_decoder_close(d);
  delete d;
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

}  // extern "C"
