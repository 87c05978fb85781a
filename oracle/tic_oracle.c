/*
 * tic_oracle.c — direct-loop C restatement of the reference codec's per-layer arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/codec_oracle.py header).  PARITY STATUS: parity unpinned
 * against TensorFlow (absent from this environment); pinned against the torch-CPU restatement and
 * the algebraic identities in tests/test_oracle.py.
 *
 * The scalar math (sigmoid, quantiser, normalise, denormalise) is the SAME source the CUDA
 * epilogues compile: include/tic_math.h.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tic_math.h"

static void same_pad(int n, int s, int* out, int* before) {
  *out = (n + s - 1) / s;
  int total = (*out - 1) * s + 3 - n;
  if (total < 0) total = 0;
  *before = total / 2;
}

/* basic_block.my_conv2d (basic_block/basic_block.py:27-47): tf.nn.conv2d SAME + bias_add + activation.
 * x [n,h,w,cin] NHWC, k HWIO [3,3,cin,cout], out [n,ho,wo,cout].  fp32 accumulation, tap-major then
 * channel order (no FMA contraction relied upon: tolerance-compared, not bit-compared). */
void tico_conv2d_same(const float* x, int n, int h, int w, int cin, const float* k, const float* b, int cout,
                      int stride, int relu, float* out) {
  int ho, wo, pt, pl;
  same_pad(h, stride, &ho, &pt);
  same_pad(w, stride, &wo, &pl);
  for (int in = 0; in < n; ++in)
    for (int oy = 0; oy < ho; ++oy)
      for (int ox = 0; ox < wo; ++ox) {
        float* o = out + (((size_t)in * ho + oy) * wo + ox) * cout;
        for (int oc = 0; oc < cout; ++oc) o[oc] = 0.0f;
        for (int kh = 0; kh < 3; ++kh) {
          int iy = oy * stride + kh - pt;
          if (iy < 0 || iy >= h) continue;
          for (int kw = 0; kw < 3; ++kw) {
            int ix = ox * stride + kw - pl;
            if (ix < 0 || ix >= w) continue;
            const float* xi = x + (((size_t)in * h + iy) * w + ix) * cin;
            const float* kk = k + (size_t)(kh * 3 + kw) * cin * cout;
            for (int ic = 0; ic < cin; ++ic) {
              float xv = xi[ic];
              const float* kr = kk + (size_t)ic * cout;
              for (int oc = 0; oc < cout; ++oc) o[oc] += xv * kr[oc];
            }
          }
        }
        for (int oc = 0; oc < cout; ++oc) {
          float v = o[oc] + b[oc];
          o[oc] = (relu && v < 0.0f) ? 0.0f : v;
        }
      }
}

/* basic_block.my_conv2d_transpose (basic_block/basic_block.py:50-71): stride 2, SAME, output 2x,
 * filter [3,3,cout,cin] (:53):  out[2i+kh, 2j+kw, oc] += x[i,j,ic] * W[kh,kw,oc,ic], indices >= 2h dropped. */
void tico_deconv2d(const float* x, int n, int h, int w, int cin, const float* k, const float* b, int cout, int relu,
                   float* out) {
  int ho = 2 * h, wo = 2 * w;
  memset(out, 0, (size_t)n * ho * wo * cout * sizeof(float));
  for (int in = 0; in < n; ++in)
    for (int i = 0; i < h; ++i)
      for (int j = 0; j < w; ++j) {
        const float* xi = x + (((size_t)in * h + i) * w + j) * cin;
        for (int kh = 0; kh < 3; ++kh) {
          int oy = 2 * i + kh;
          if (oy >= ho) continue;
          for (int kw = 0; kw < 3; ++kw) {
            int ox = 2 * j + kw;
            if (ox >= wo) continue;
            float* o = out + (((size_t)in * ho + oy) * wo + ox) * cout;
            const float* kk = k + (size_t)(kh * 3 + kw) * cout * cin;
            for (int oc = 0; oc < cout; ++oc) {
              const float* kr = kk + (size_t)oc * cin;
              float acc = 0.0f;
              for (int ic = 0; ic < cin; ++ic) acc += xi[ic] * kr[ic];
              o[oc] += acc;
            }
          }
        }
      }
  size_t total = (size_t)n * ho * wo;
  for (size_t p = 0; p < total; ++p)
    for (int oc = 0; oc < cout; ++oc) {
      float v = out[p * cout + oc] + b[oc];
      out[p * cout + oc] = (relu && v < 0.0f) ? 0.0f : v;
    }
}

void tico_sigmoid(const float* x, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = tic_sigmoid_f32(x[i]);
}

/* model_0/model.py:137-138 */
void tico_quantize(const float* x, uint8_t* out, int64_t n, int quan_scale) {
  for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)tic_quantize_symbol(x[i], quan_scale);
}

/* model_0/model.py:44; x is [npix,3] */
void tico_normalize(const float* x, float* out, int64_t npix, const float* mean, const float* stdv) {
  for (int64_t i = 0; i < npix; ++i)
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = tic_normalize(x[i * 3 + c], mean[c], stdv[c]);
}

/* model_0/model.py:251,259; y is [npix,3] */
void tico_denorm_clip(const float* y, float* out, int64_t npix, const float* mean, const float* stdv) {
  for (int64_t i = 0; i < npix; ++i)
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = tic_denorm_clip(y[i * 3 + c], mean[c], stdv[c]);
}
