/*
 * tic_math.h — single-source fp32 scalar math shared by the CUDA epilogues
 * (tf_image_compression_b200/csrc) and the CPU oracle (oracle/tic_oracle.c).
 *
 * Why single-source: the quantiser of the reference
 *     output = tf.nn.sigmoid(output) * (quan_scale - 1)
 *     output = tf.stop_gradient(tf.round(output) - output) + output
 * (model_0/model.py:137-138) decides a symbol on the last ulp of sigmoid() when
 * the logit is close to a rounding boundary.  With q = 2 every logit in
 * (0, ~1.2e-7] gives fl32(sigmoid) == 0.5 and rounds (half-to-even) to 0.  The
 * boundary is therefore a property of the sigmoid implementation; compiling the
 * SAME explicit-fmaf code with gcc and nvcc makes it bit-identical on both
 * sides.  No fast-math, no contraction: every operation below is an explicit
 * IEEE-754 binary32 add/mul/div or a fused multiply-add.
 *
 * TensorFlow's own sigmoid (Eigen) is not available in this environment, so
 * agreement with TF in the last ulp of sigmoid is unpinned.
 */
#ifndef TIC_MATH_H_
#define TIC_MATH_H_

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TIC_HD __host__ __device__ __forceinline__
#else
#define TIC_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define TIC_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#define TIC_MULF(a, b) __fmul_rn((a), (b))
#define TIC_ADDF(a, b) __fadd_rn((a), (b))
#define TIC_DIVF(a, b) __fdiv_rn((a), (b))
#define TIC_RINTF(a) rintf(a)
#else
#include <math.h>
#define TIC_FMAF(a, b, c) fmaf((a), (b), (c))
/* volatile stops the host compiler from contracting mul+add into an fma */
TIC_HD float tic_mulf_(float a, float b) { volatile float r = a * b; return r; }
TIC_HD float tic_addf_(float a, float b) { volatile float r = a + b; return r; }
TIC_HD float tic_divf_(float a, float b) { volatile float r = a / b; return r; }
#define TIC_MULF(a, b) tic_mulf_((a), (b))
#define TIC_ADDF(a, b) tic_addf_((a), (b))
#define TIC_DIVF(a, b) tic_divf_((a), (b))
#define TIC_RINTF(a) rintf(a) /* default rounding mode: nearest-even */
#endif

TIC_HD float tic_bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

/* exp(x) for x clamped to [-87, 87]: Cody-Waite reduction by ln2 (hi/lo), degree-6
 * polynomial on [-ln2/2, ln2/2] evaluated with fma (Horner), exact scale by 2^n. */
TIC_HD float tic_expf(float x) {
  x = x < -87.0f ? -87.0f : x;
  x = x > 87.0f ? 87.0f : x;
  const float kLog2e = 1.44269504088896341f;
  const float kLn2Hi = 0.693145751953125f;       /* 0x3f317200 */
  const float kLn2Lo = 1.42860682030941723e-6f;  /* ln2 - kLn2Hi */
  float n = TIC_RINTF(TIC_MULF(x, kLog2e));
  float r = TIC_FMAF(n, -kLn2Hi, x);
  r = TIC_FMAF(n, -kLn2Lo, r);
  float p = 1.0f / 720.0f;
  p = TIC_FMAF(p, r, 1.0f / 120.0f);
  p = TIC_FMAF(p, r, 1.0f / 24.0f);
  p = TIC_FMAF(p, r, 1.0f / 6.0f);
  p = TIC_FMAF(p, r, 0.5f);
  p = TIC_FMAF(p, r, 1.0f);
  p = TIC_FMAF(p, r, 1.0f);
  int32_t e = (int32_t)n; /* |n| <= 126 after the clamp */
  float s = tic_bits_to_float((uint32_t)(e + 127) << 23);
  return TIC_MULF(p, s);
}

/* sigmoid(x) = 1 / (1 + exp(-x)); follows tf.nn.sigmoid at model_0/model.py:137 */
TIC_HD float tic_sigmoid_f32(float x) {
  return TIC_DIVF(1.0f, TIC_ADDF(1.0f, tic_expf(-x)));
}

/* Bottleneck quantiser, model_0/model.py:137-138:
 *   o = sigmoid(x) * (q - 1);  value = (round(o) - o) + o  == round(o) in fp32
 * tf.round is round-half-to-even == rintf.  Returns the integer symbol in [0, q-1]. */
TIC_HD int32_t tic_quantize_symbol(float logit, int32_t quan_scale) {
  float o = TIC_MULF(tic_sigmoid_f32(logit), (float)(quan_scale - 1));
  return (int32_t)TIC_RINTF(o);
}

/* (x - mean) / std : model_0/model.py:44, true division (not reciprocal multiply) */
TIC_HD float tic_normalize(float x, float mean, float stdv) {
  return TIC_DIVF(TIC_ADDF(x, -mean), stdv);
}

/* clip(y * std + mean, 0, 255) : model_0/model.py:251,259 (Mul then Add, two roundings) */
TIC_HD float tic_denorm_clip(float y, float mean, float stdv) {
  float v = TIC_ADDF(TIC_MULF(y, stdv), mean);
  v = v < 0.0f ? 0.0f : v;
  v = v > 255.0f ? 255.0f : v;
  return v;
}

#endif /* TIC_MATH_H_ */
