/*
 * NumPy's float32 exp as computed by its AVX2 / AVX-512F SIMD loop
 * (the value `np.exp(float32_array)` returns on any AVX2-or-better host), restated
 * as scalar code: Cody-Waite range reduction with FMA, a degree-5 / degree-2
 * rational, an IEEE divide and an exact power-of-two scale.  This is what
 * ExponentialBoosting.process evaluates (bithtm/regularizations.py:16); the boost
 * factor must match it bit for bit or the boosted overlaps differ.
 *
 * Usable from C (gcc -mfma -ffp-contract=off), C++ and CUDA device code.  Every
 * operation is individually rounded; only the explicit fma calls are fused.
 * Valid for -87 < x <= 0 (normal results); boosting only produces x <= 0.
 */
#ifndef BH_NP_EXPF_H
#define BH_NP_EXPF_H

#if defined(__CUDA_ARCH__)
#define BH_EXP_FN __device__ __forceinline__
#define BH_MUL(a, b) __fmul_rn((a), (b))
#define BH_ADD(a, b) __fadd_rn((a), (b))
#define BH_SUB(a, b) __fsub_rn((a), (b))
#define BH_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define BH_DIV(a, b) __fdiv_rn((a), (b))
#define BH_F2I(f) __float_as_int(f)
#define BH_I2F(i) __int_as_float(i)
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define BH_EXP_FN static inline
#define BH_MUL(a, b) ((a) * (b))
#define BH_ADD(a, b) ((a) + (b))
#define BH_SUB(a, b) ((a) - (b))
#define BH_FMA(a, b, c) fmaf((a), (b), (c))
#define BH_DIV(a, b) ((a) / (b))
static inline int32_t bh_f2i_(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float bh_i2f_(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
#define BH_F2I(f) bh_f2i_(f)
#define BH_I2F(i) bh_i2f_(i)
#endif

BH_EXP_FN float bh_np_expf(float x) {
  const float log2e = 1.44269504088896340736f;
  const float magic = 12582912.0f; /* 0x1.8p+23: round-to-nearest-integer trick */
  const float c1 = -6.93145752e-1f, c2 = -1.42860677e-6f; /* -ln2 split hi/lo */
  const float p0 = 9.999999999980870924916e-01f, p1 = 7.257664613233124478488e-01f,
              p2 = 2.473615434895520810817e-01f, p3 = 5.114512081637298353406e-02f,
              p4 = 6.757896990527504603057e-03f, p5 = 5.082762527590693718096e-04f;
  const float q0 = 1.0f, q1 = -2.742335390411667452936e-01f, q2 = 2.159509375685829852307e-02f;

  float q = BH_MUL(x, log2e);
  q = BH_ADD(q, magic);
  q = BH_SUB(q, magic);
  float r = BH_FMA(q, c1, x);
  r = BH_FMA(q, c2, r);
  float num = BH_FMA(p5, r, p4);
  num = BH_FMA(num, r, p3);
  num = BH_FMA(num, r, p2);
  num = BH_FMA(num, r, p1);
  num = BH_FMA(num, r, p0);
  float den = BH_FMA(q2, r, q1);
  den = BH_FMA(den, r, q0);
  float y = BH_DIV(num, den);
  /* y * 2^q, exact while the result is normal: add q to the exponent field */
  int qi = (int)q;
  return BH_I2F(BH_F2I(y) + qi * 8388608);
}

#endif /* BH_NP_EXPF_H */
