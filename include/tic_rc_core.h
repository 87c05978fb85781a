/*
 * tic_rc_core.h — the arithmetic of the entropy coder, single source for the host library
 * (tf_image_compression_b200/csrc/rangecoder.cpp, behind include/tic_rangecoder.h) and the CUDA entropy stage
 * (tf_image_compression_b200/csrc/tic_entropy.cuh, behind tic_entropy_* in include/tic.h): both include this file, so
 * "GPU bitstreams byte-identical to the host coder" holds by construction and is checked byte for byte in the tests.
 *
 * It stands where the third-party PyPI package `range_coder` stands in the reference (encode.py:76-97,
 * decode.py:79-101).  That package is not vendored and not installable here, so this is a restatement of the
 * published carry-propagating byte-wise range coder (G. N. N. Martin 1979; the low / range / cache / pending-0xFF
 * formulation used by LZMA), pinned by the reference's own known-answer test (other/test_range_coder.py:37-68).
 * Byte identity with the PyPI package beyond that test is UNPINNED (its first four stream bytes differ by its own
 * test's admission, "the first 4 bytes are special").
 *
 * Stream format ("tic range coder v2")
 *   - 32-bit range, 33-bit low.  Initially low = 0, range = 2^24 (= probability 1).  Per symbol with cumulative
 *     frequencies [lo, hi) of `total` (total <= 2^16):  r = range / total;  low += r * lo;  range = r * (hi - lo);
 *     while range < 2^24: ship the top byte of low (carry-propagating), low <<= 8, range <<= 8.
 *   - The first two shipped bytes are always zero (integer part at the initial scale) and are not stored.
 *   - close(): low is moved to the value inside [low, low + range) with the most trailing zero bytes, five bytes are
 *     shipped, and trailing 0x00 bytes of the whole stream are not stored: the decoder reads zeros past the end.
 *     (A dyadic source that ends on a byte boundary therefore costs exactly its entropy: the reference's
 *     known-answer vector is 17 bytes of 0x0b.)
 *   - Several encode() calls with different tables may share one stream (the state just carries on).
 *   - SEGMENTED CALLS.  A range coder is a serial recurrence: one image of BASELINE config 2 is 786 432 symbols, and 64
 *     such streams keep neither 32 host threads nor a GPU busy.  So an encode() call of MORE than TIC_RC_SEGMENT_SYMBOLS
 *     symbols made on a fresh stream (nothing coded yet, or right after another segmented call) is written as a
 *     container: nseg = ceil(n / TIC_RC_SEGMENT_SYMBOLS) little-endian uint32 byte counts, then nseg independent streams
 *     (each coded from the initial state, terminated and zero-stripped as above) of TIC_RC_SEGMENT_SYMBOLS symbols each
 *     (the last one shorter).  After it the stream is fresh again.  The decoder applies the same rule to its decode(n)
 *     calls, so nothing is stored beyond the byte counts.  Calls of at most TIC_RC_SEGMENT_SYMBOLS symbols (every test of
 *     the reference's range_coder suite, including its known-answer vector) are plain streams.  Overhead: 4 bytes and one
 *     termination (<= 5 bytes) per 32 768 symbols.
 * The range recurrence does not depend on low, and a binary symbol costs one shift, one multiply and one compare on
 * the critical path; the segments are what both the host thread pool and the GPU stage parallelise over.
 */
#ifndef TIC_RC_CORE_H_
#define TIC_RC_CORE_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define TIC_RC_HD __host__ __device__ __forceinline__
#else
#define TIC_RC_HD inline
#endif

#define TIC_RC_TOP (1u << 24)          /* renormalise below this */
#define TIC_RC_MAX_TOTAL (1u << 16)    /* largest frequency total: r = range / total stays >= 256 */

#define TIC_RC_SEGMENT_SYMBOLS 32768    /* calls longer than this on a fresh stream are coded as independent segments */

/* Worst-case stored bytes of a plain stream of n symbols (every symbol at the smallest width of the largest total costs
 * 16.006 bits) plus the flush ... */
TIC_RC_HD int64_t tic_rc_plain_bound(int64_t n) { return 2 * n + (n >> 6) + 16; }
TIC_RC_HD int64_t tic_rc_segments(int64_t n) {
  return n > TIC_RC_SEGMENT_SYMBOLS ? (n + TIC_RC_SEGMENT_SYMBOLS - 1) / TIC_RC_SEGMENT_SYMBOLS : 0;
}
/* ... and of what one encode() call of n symbols on a fresh stream writes (container header included) */
TIC_RC_HD int64_t tic_rc_bound(int64_t n) {
  const int64_t nseg = tic_rc_segments(n);
  return nseg ? nseg * (4 + tic_rc_plain_bound(TIC_RC_SEGMENT_SYMBOLS)) : tic_rc_plain_bound(n);
}

typedef struct tic_rc_enc_state {
  uint64_t low;
  uint32_t range;
  uint32_t cache;
  uint32_t pending;  /* LZMA's cacheSize: 1 + number of 0xFF bytes waiting for a possible carry */
  uint32_t skip;     /* leading shipped bytes still to drop (2 at the start) */
} tic_rc_enc_state;

TIC_RC_HD void tic_rc_enc_init(tic_rc_enc_state* s) {
  s->low = 0;
  s->range = TIC_RC_TOP;
  s->cache = 0;
  s->pending = 1;
  s->skip = 2;
}

/* Sink: any type with  void put(uint8_t)  */
template <class Sink>
TIC_RC_HD void tic_rc_ship(tic_rc_enc_state* s, Sink& out, uint32_t byte) {
  if (s->skip) {
    --s->skip;
    return;
  }
  out.put((uint8_t)byte);
}

template <class Sink>
TIC_RC_HD void tic_rc_shift_low(tic_rc_enc_state* s, Sink& out) {
  const uint32_t lo32 = (uint32_t)s->low;
  const uint32_t carry = (uint32_t)(s->low >> 32);
  if (lo32 < 0xFF000000u || carry != 0) {
    uint32_t b = s->cache;
    do {
      tic_rc_ship(s, out, (b + carry) & 0xFFu);
      b = 0xFFu;
    } while (--s->pending != 0);
    s->cache = lo32 >> 24;
  }
  ++s->pending;
  s->low = (uint64_t)((lo32 & 0x00FFFFFFu) << 8);
}

/* one symbol with cumulative frequencies [lo, hi) out of total; r = range / total supplied by the caller
 * (a shift when total is a power of two) */
template <class Sink>
TIC_RC_HD void tic_rc_enc_step(tic_rc_enc_state* s, Sink& out, uint32_t r, uint32_t lo, uint32_t hi) {
  s->low += (uint64_t)r * lo;
  s->range = r * (hi - lo);
  while (s->range < TIC_RC_TOP) {
    tic_rc_shift_low(s, out);
    s->range <<= 8;
  }
}

/* binary alphabet, total = 2^k (every shipped config: quan_scale = 2, resolution = 4096): cum = [0, c1, 2^k].
 * Identical arithmetic to tic_rc_enc_step with r = range >> k; r * (2^k - c1) is (r << k) - r * c1: one multiply. */
template <class Sink>
TIC_RC_HD void tic_rc_enc_bit(tic_rc_enc_state* s, Sink& out, uint32_t bit, int k, uint32_t c1) {
  const uint32_t r = s->range >> k;
  const uint32_t t = r * c1;
  const uint32_t m = 0u - bit;   /* branch-free select: the symbols are close to coin flips */
  s->low += (uint64_t)(t & m);
  s->range = (t & ~m) | (((r << k) - t) & m);
  while (s->range < TIC_RC_TOP) {
    tic_rc_shift_low(s, out);
    s->range <<= 8;
  }
}

template <class Sink>
TIC_RC_HD void tic_rc_enc_finish(tic_rc_enc_state* s, Sink& out) {
  /* the value in [low, low + range - 1] with the most trailing zero bytes */
  const uint64_t last = s->low + (uint64_t)s->range - 1;
  for (int k = 4; k >= 0; --k) {
    const uint64_t unit = (uint64_t)1 << (8 * k);
    const uint64_t v = (s->low + unit - 1) / unit * unit;
    if (v <= last) {
      s->low = v;
      break;
    }
  }
  for (int i = 0; i < 5; ++i) tic_rc_shift_low(s, out);
  /* the sink drops the trailing zero bytes of the stream */
}

typedef struct tic_rc_dec_state {
  uint32_t range;
  uint32_t code;
  int primed;
} tic_rc_dec_state;

TIC_RC_HD void tic_rc_dec_init(tic_rc_dec_state* s) {
  s->range = TIC_RC_TOP;
  s->code = 0;
  s->primed = 0;
}

/* Source: any type with  uint32_t get()  returning the next stored byte, 0 past the end */
template <class Source>
TIC_RC_HD void tic_rc_dec_prime(tic_rc_dec_state* s, Source& in) {
  if (!s->primed) {
    s->code = in.get() << 16;
    s->code |= in.get() << 8;
    s->code |= in.get();
    s->primed = 1;
  }
}

/* position of the code value in the table: v in [0, total) (clamped for corrupt streams); r = range / total */
TIC_RC_HD uint32_t tic_rc_dec_target(const tic_rc_dec_state* s, uint32_t r, uint32_t total) {
  const uint32_t v = s->code / r;
  return v < total ? v : total - 1;
}

template <class Source>
TIC_RC_HD void tic_rc_dec_step(tic_rc_dec_state* s, Source& in, uint32_t r, uint32_t lo, uint32_t hi) {
  s->code -= r * lo;
  s->range = r * (hi - lo);
  while (s->range < TIC_RC_TOP) {
    s->code = (s->code << 8) | in.get();
    s->range <<= 8;
  }
}

#endif /* TIC_RC_CORE_H_ */
