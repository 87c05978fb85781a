// GPU entropy stage: the per-image static-model range coder of encode.py:171-202 /
// decode.py:182-208 on the device, bit-exact with the host coder (librangecoder.so) because both compile the same
// arithmetic (include/tic_rc_core.h).  A stream (one image's symbols, patch-major then h, w, c: encode.py:171-182) is a
// serial recurrence, so the unit of parallelism is the stream: one warp per stream, lane 0 runs the coder (the other
// lanes only help load the table), one stream per SM sub-partition scheduler so every stream issues at full
// single-thread rate.  With the binary alphabet of every shipped config (quan_scale = 2) and the power-of-two table
// total the reference uses (resolution = 4096, encode.py:91) a symbol costs one shift, one multiply, one compare and
// an occasional byte store on the critical path.  Symbols stay in HBM after the encoder's last layer and only the
// compressed bytes cross PCIe (0.25 sym/px -> ~0.03 B/px instead of 0.25 B/px for model_0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tic_rc_core.h"

namespace tic {

constexpr int kEntropyMaxSymbols = 256;

struct DevSink {
  uint8_t* p;
  long long cap, pos, len;
  int overflow;
  __device__ __forceinline__ void put(uint8_t b) {
    if (pos >= cap) {
      overflow = 1;
      return;
    }
    p[pos++] = b;
    if (b) len = pos;
  }
};

// Stored bytes through 16-byte loads, the block after the current one already in flight; zeros past the end.
struct DevSource {
  const uint4* p;     // 16-byte aligned
  long long nbytes;   // stored length
  long long blk;      // index of the block in `cur`
  uint4 cur, nxt;
  uint32_t off;       // byte offset inside cur
  __device__ __forceinline__ uint4 load(long long b) const {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const long long first = b * 16;
    if (first < nbytes) {
      v = __ldg(p + b);
      const long long left = nbytes - first;
      if (left < 16) {  // clear the bytes past the stored length (the slot behind them is not the stream's)
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long lo = 4 * i;
          if (left <= lo) w[i] = 0u;
          else if (left < lo + 4) w[i] &= (1u << (8 * (int)(left - lo))) - 1u;
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    return v;
  }
  __device__ __forceinline__ void init(const uint8_t* base, long long n) {
    p = reinterpret_cast<const uint4*>(base);
    nbytes = n;
    blk = 0;
    off = 0;
    cur = load(0);
    nxt = load(1);
  }
  __device__ __forceinline__ uint32_t get() {
    if (off == 16) {
      cur = nxt;
      ++blk;
      nxt = load(blk + 1);
      off = 0;
    }
    const uint32_t w = off < 8 ? (off < 4 ? cur.x : cur.y) : (off < 12 ? cur.z : cur.w);
    const uint32_t b = (w >> (8 * (off & 3))) & 0xFFu;
    ++off;
    return b;
  }
};

// status codes written to *status (first error wins): 1 = symbol outside the table / of zero probability,
// 2 = output slot too small
__global__ void __launch_bounds__(32) rc_encode_kernel(const uint8_t* __restrict__ sym, long long stream_len, const uint32_t* __restrict__ cum,
                                                       int n_cum, uint8_t* __restrict__ out, long long out_stride,
                                                       long long* __restrict__ out_bytes, unsigned int* status) {
  __shared__ uint32_t s_cum[kEntropyMaxSymbols + 1];
  for (int i = threadIdx.x; i < n_cum; i += 32) s_cum[i] = cum[i];
  __syncwarp();
  if (threadIdx.x != 0) return;
  const long long stream = blockIdx.x;
  const uint8_t* s = sym + stream * stream_len;
  DevSink sink{out + stream * out_stride, out_stride, 0, 0, 0};
  tic_rc_enc_state st;
  tic_rc_enc_init(&st);
  const uint32_t total = s_cum[n_cum - 1];
  const bool p2 = (total & (total - 1)) == 0;
  const int k = 31 - __clz(total);
  const int nsym = n_cum - 1;
  int bad = 0;
  // 16 symbols per load when the stream is 16-byte aligned; the next block is requested before this one is coded
  const bool aligned = ((reinterpret_cast<uintptr_t>(s) | (uintptr_t)stream_len) & 15) == 0;
  if (aligned && nsym == 2 && p2) {
    const uint32_t c1 = s_cum[1], w1 = total - c1;
    const uint4* s16 = reinterpret_cast<const uint4*>(s);
    const long long nblk = stream_len >> 4;
    uint4 nxt = nblk ? __ldg(s16) : make_uint4(0u, 0u, 0u, 0u);
    // The symbol loop is deliberately NOT unrolled: one warp runs alone on its scheduler, so what counts is the dependent
    // chain of a symbol (shift, multiply, select, compare) and an instruction footprint that stays inside the
    // instruction cache (the fully unrolled 16-symbol body was 66 KB of SASS and ran at ~95 cycles per symbol).
    for (long long b = 0; b < nblk; ++b) {
      uint4 cur = nxt;
      if (b + 1 < nblk) nxt = __ldg(s16 + b + 1);
      // symbols other than 0 / 1, or of zero width, are an error (checked once per block, outside the critical path)
      const uint32_t any = cur.x | cur.y | cur.z | cur.w, all = cur.x & cur.y & cur.z & cur.w;
      if (any & 0xFEFEFEFEu) bad = 1;
      if ((any && w1 == 0) || ((~all & 0x01010101u) && c1 == 0)) bad = 1;
      if (bad) break;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        uint32_t word = cur.x;
        cur.x = cur.y;
        cur.y = cur.z;
        cur.z = cur.w;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          tic_rc_enc_bit(&st, sink, word & 1u, k, c1);
          word >>= 8;
        }
      }
    }
  } else {
    for (long long i = 0; i < stream_len; ++i) {
      const int v = s[i];
      if (v >= nsym) {
        bad = 1;
        break;
      }
      const uint32_t lo = s_cum[v], hi = s_cum[v + 1];
      if (hi == lo) {
        bad = 1;
        break;
      }
      const uint32_t r = p2 ? st.range >> k : st.range / total;
      tic_rc_enc_step(&st, sink, r, lo, hi);
    }
  }
  if (!bad) tic_rc_enc_finish(&st, sink);
  out_bytes[stream] = bad ? 0 : sink.len;
  if (bad) atomicCAS(status, 0u, 1u);
  else if (sink.overflow) atomicCAS(status, 0u, 2u);
}

__global__ void __launch_bounds__(32) rc_decode_kernel(const uint8_t* __restrict__ in, long long in_stride, const long long* __restrict__ in_bytes,
                                                       const uint32_t* __restrict__ cum, int n_cum, uint8_t* __restrict__ sym,
                                                       long long stream_len) {
  __shared__ uint32_t s_cum[kEntropyMaxSymbols + 1];
  for (int i = threadIdx.x; i < n_cum; i += 32) s_cum[i] = cum[i];
  __syncwarp();
  if (threadIdx.x != 0) return;
  const long long stream = blockIdx.x;
  DevSource src;
  const long long stored = in_bytes[stream];
  src.init(in + stream * in_stride, stored < in_stride ? stored : in_stride);
  uint8_t* o = sym + stream * stream_len;
  tic_rc_dec_state st;
  tic_rc_dec_init(&st);
  tic_rc_dec_prime(&st, src);
  const uint32_t total = s_cum[n_cum - 1];
  const bool p2 = (total & (total - 1)) == 0;
  const int k = 31 - __clz(total);
  const bool aligned = ((reinterpret_cast<uintptr_t>(o) | (uintptr_t)stream_len) & 15) == 0;
  if (aligned && n_cum == 3 && p2) {
    const uint32_t c1 = s_cum[1];
    uint4* o16 = reinterpret_cast<uint4*>(o);
    for (long long b = 0; b < (stream_len >> 4); ++b) {
      uint4 o4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        uint32_t word = 0u;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const uint32_t r = st.range >> k;
          const uint32_t t = r * c1;
          const bool one = c1 == total ? false : st.code >= t;
          word |= (one ? 1u : 0u) << (8 * j);
          tic_rc_dec_step(&st, src, r, one ? c1 : 0u, one ? total : c1);
        }
        o4.x = o4.y;
        o4.y = o4.z;
        o4.z = o4.w;
        o4.w = word;
      }
      o16[b] = o4;
    }
  } else {
    for (long long i = 0; i < stream_len; ++i) {
      const uint32_t r = p2 ? st.range >> k : st.range / total;
      const uint32_t v = tic_rc_dec_target(&st, r, total);
      // last entry with cum[s] <= v (skips zero-width symbols): binary search over the shared table
      int lo = 0, hi = n_cum - 1;  // invariant: cum[lo] <= v < cum[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_cum[mid] <= v) lo = mid;
        else hi = mid;
      }
      // step over zero-width entries that share cum[lo] (upper_bound semantics: the LAST entry <= v)
      o[i] = (uint8_t)lo;
      tic_rc_dec_step(&st, src, r, s_cum[lo], s_cum[lo + 1]);
    }
  }
}

}  // namespace tic
