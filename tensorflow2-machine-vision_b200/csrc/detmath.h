// detmath.h — deterministic fp32 elementary functions for the detection hot path.
//
// Every function here is built only from IEEE-754 binary32 add / sub / mul / div / sqrt
// (each individually rounded to nearest-even, never contracted into an FMA) plus integer
// bit manipulation.  The same source therefore produces bit-identical results when compiled
//   * by nvcc for sm_100a   (ops spelled __fadd_rn/__fmul_rn/... so ptxas cannot fuse them), and
//   * by gcc for the host   (-ffp-contract=off, SSE2 scalar fp32).
// That is what lets discrete outputs (NMS keep order, anchor argmax, thresholds on sigmoid
// scores) be compared bit-for-bit between the CUDA kernels and the CPU oracle, while the values
// themselves stay within a few ulp of the correctly rounded result (see oracle/verify_detmath.c,
// which sweeps all 2^32 inputs against double-precision libm; summary in oracle/DETMATH_REPORT.md).
//
// The reference evaluates these through TensorFlow/Eigen ops (tf.sigmoid, tf.exp, tf.math.log,
// tf.math.atan, pow):  utils/tf_yolo_utils.py:57,61,139-140,153-155; utils/tf_iou_utils.py:50,55;
// efficientnet/utils/anchors.py:241-242,266-267; losses/focal_loss.py:40-47.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define DM_ADD(a, b) __fadd_rn((a), (b))
#define DM_SUB(a, b) __fsub_rn((a), (b))
#define DM_MUL(a, b) __fmul_rn((a), (b))
#define DM_DIV(a, b) __fdiv_rn((a), (b))
#define DM_SQRT(a) __fsqrt_rn((a))
#else
#define DM_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define DM_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define DM_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define DM_DIV(a, b) ((float)((float)(a) / (float)(b)))
#define DM_SQRT(a) (sqrtf((float)(a)))
#endif

DM_HD uint32_t dm_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
DM_HD float dm_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

#define DM_INF_BITS 0x7f800000u
#define DM_NAN_BITS 0x7fc00000u

DM_HD float dm_fabsf(float x) { return dm_u2f(dm_f2u(x) & 0x7fffffffu); }
DM_HD int dm_isnan(float x) { return (dm_f2u(x) & 0x7fffffffu) > DM_INF_BITS; }
DM_HD int dm_isinf(float x) { return (dm_f2u(x) & 0x7fffffffu) == DM_INF_BITS; }
// max/min with "first operand wins unless the second is strictly better"; NaN in either
// operand is outside the parity contract (DESIGN.md §non-finite inputs).
DM_HD float dm_max(float a, float b) { return (b > a) ? b : a; }
DM_HD float dm_min(float a, float b) { return (b < a) ? b : a; }

// 2^k for k in [-126, 127] (normal range).
DM_HD float dm_pow2i(int k) { return dm_u2f((uint32_t)(k + 127) << 23); }

// ---------------------------------------------------------------------------------------------
// exp(x).  k = rint(x*log2(e)); r = x - k*ln2 (two-constant Cody-Waite, k*LN2_HI exact);
// e^r by a degree-7 Taylor polynomial on |r| <= 0.3466 (truncation < 5e-9 relative);
// result scaled by 2^k in two exact/once-rounded steps so denormal results round once.
DM_HD float dm_expf(float x) {
  if (dm_isnan(x)) return x;
  if (x > 88.72283935546875f) return dm_u2f(DM_INF_BITS);
  if (x < -103.97208404541016f) return 0.0f;
  float t = DM_MUL(x, 1.44269502162933349609375f);
  float kf = DM_SUB(DM_ADD(t, 12582912.0f), 12582912.0f);
  float r = DM_SUB(DM_SUB(x, DM_MUL(kf, 0.693145751953125f)), DM_MUL(kf, 1.428606765330187045037746429443359375e-06f));
  float q = 1.98412701138295233249664306640625e-04f;           // 1/5040
  q = DM_ADD(DM_MUL(q, r), 1.38888892251998186111450195312e-03f);  // 1/720
  q = DM_ADD(DM_MUL(q, r), 8.3333337679505348205566406250e-03f);   // 1/120
  q = DM_ADD(DM_MUL(q, r), 4.16666679084300994873046875e-02f);     // 1/24
  q = DM_ADD(DM_MUL(q, r), 1.666666716337203979492187500e-01f);    // 1/6
  q = DM_ADD(DM_MUL(q, r), 0.5f);
  float y = DM_ADD(1.0f, DM_ADD(r, DM_MUL(DM_MUL(r, r), q)));
  int k = (int)kf;
  int k1 = k >> 1;  // floor(k/2)
  int k2 = k - k1;
  return DM_MUL(DM_MUL(y, dm_pow2i(k1)), dm_pow2i(k2));
}

// sigmoid(x) = 1/(1+e^-x) for x >= 0, e^x/(1+e^x) for x < 0 (both branches avoid cancellation).
DM_HD float dm_sigmoidf(float x) {
  if (dm_isnan(x)) return x;
  if (x >= 0.0f) {
    float e = dm_expf(-x);
    return DM_DIV(1.0f, DM_ADD(1.0f, e));
  }
  float e = dm_expf(x);
  return DM_DIV(e, DM_ADD(1.0f, e));
}

// ---------------------------------------------------------------------------------------------
// log(x).  x = m*2^e with m in [sqrt(1/2), sqrt(2)); f = m-1 (exact); s = f/(2+f);
// log(1+f) = f - f^2/2 + s*(f^2/2 + R(s^2)), R = 2/3 s^2 + 2/5 s^4 + 2/7 s^6 + 2/9 s^8 with the
// classic minimax-adjusted constants; e*ln2 added as a hi/lo pair (e*LN2_HI exact).
DM_HD float dm_logf(float x) {
  uint32_t ix = dm_f2u(x);
  if (dm_isnan(x)) return x;
  if ((ix & 0x7fffffffu) == 0u) return dm_u2f(0xff800000u);  // log(+-0) = -inf
  if (ix & 0x80000000u) return dm_u2f(DM_NAN_BITS);            // log(negative) = NaN
  if (ix == DM_INF_BITS) return x;
  int e = 0;
  if (ix < 0x00800000u) {  // subnormal: scale by 2^25 (exact)
    x = DM_MUL(x, 33554432.0f);
    ix = dm_f2u(x);
    e = -25;
  }
  e += (int)(ix >> 23) - 127;
  uint32_t im = (ix & 0x007fffffu) | 0x3f800000u;  // m in [1,2)
  if (im > 0x3fb504f3u) {                           // m > sqrt(2): halve (exact)
    im -= 0x00800000u;
    e += 1;
  }
  float m = dm_u2f(im);
  float f = DM_SUB(m, 1.0f);
  float s = DM_DIV(f, DM_ADD(2.0f, f));
  float z = DM_MUL(s, s);
  float R = 0.24279078841209411621f;
  R = DM_ADD(DM_MUL(R, z), 0.28498786687850952148f);
  R = DM_ADD(DM_MUL(R, z), 0.40000972151756286621f);
  R = DM_ADD(DM_MUL(R, z), 0.66666662693023681640625f);
  R = DM_MUL(R, z);
  float hfsq = DM_MUL(0.5f, DM_MUL(f, f));
  float ef = (float)e;
  float lo = DM_ADD(DM_MUL(s, DM_ADD(hfsq, R)), DM_MUL(ef, 9.0580006144591607153415679931640625e-06f));
  // e*ln2_hi - ((hfsq - lo) - f)
  return DM_SUB(DM_MUL(ef, 0.693138122558593750f), DM_SUB(DM_SUB(hfsq, lo), f));
}

// log1p(y) in the compensated form Eigen uses (numext::log1p): log(u)*(y/(u-1)), u = 1+y.
DM_HD float dm_log1pf(float y) {
  float u = DM_ADD(1.0f, y);
  if (u == 1.0f) return y;
  if (dm_isinf(u) || dm_isnan(u)) return u;
  return DM_MUL(dm_logf(u), DM_DIV(y, DM_SUB(u, 1.0f)));
}

// ---------------------------------------------------------------------------------------------
// atan(x): argument reduction at 7/16, 11/16, 19/16, 39/16 onto atan(1/2), atan(1), atan(3/2),
// pi/2 (hi/lo pairs) and a 5-term odd minimax polynomial on |t| < 7/16.
DM_HD float dm_atanf(float x) {
  if (dm_isnan(x)) return x;
  uint32_t sign = dm_f2u(x) & 0x80000000u;
  float ax = dm_fabsf(x);
  float res;
  if (ax >= 67108864.0f) {  // 2^26: atan = pi/2 to fp32
    res = 1.57079637050628662109375f;
  } else {
    int id;
    float t;
    if (ax < 0.4375f) {
      if (ax < 2.44140625e-4f) return x;  // |x| < 2^-12: atan(x) = x to fp32
      id = -1;
      t = ax;
    } else if (ax < 0.6875f) {
      id = 0;
      t = DM_DIV(DM_SUB(DM_MUL(2.0f, ax), 1.0f), DM_ADD(2.0f, ax));
    } else if (ax < 1.1875f) {
      id = 1;
      t = DM_DIV(DM_SUB(ax, 1.0f), DM_ADD(ax, 1.0f));
    } else if (ax < 2.4375f) {
      id = 2;
      t = DM_DIV(DM_SUB(ax, 1.5f), DM_ADD(1.0f, DM_MUL(1.5f, ax)));
    } else {
      id = 3;
      t = DM_DIV(-1.0f, ax);
    }
    float z = DM_MUL(t, t);
    float w = DM_MUL(z, z);
    float s1 = DM_MUL(z, DM_ADD(3.3333328366e-01f, DM_MUL(w, DM_ADD(1.4253635705e-01f, DM_MUL(w, 6.1687607318e-02f)))));
    float s2 = DM_MUL(w, DM_ADD(-1.9999158382e-01f, DM_MUL(w, -1.0648017377e-01f)));
    float ts = DM_MUL(t, DM_ADD(s1, s2));
    if (id < 0) {
      res = DM_SUB(t, ts);
    } else {
      float hi, lo;
      if (id == 0) { hi = 4.6364760399e-01f; lo = 5.0121582440e-09f; }
      else if (id == 1) { hi = 7.8539812565e-01f; lo = 3.7748947079e-08f; }
      else if (id == 2) { hi = 9.8279368877e-01f; lo = 3.4473217170e-08f; }
      else { hi = 1.5707962513e+00f; lo = 7.5497894159e-08f; }
      res = DM_SUB(hi, DM_SUB(DM_SUB(ts, lo), t));
    }
  }
  return dm_u2f(dm_f2u(res) | sign);
}

// pow(x, y) for the two call sites on the path: d^0.6 with d in [0,1] (tf_iou_utils.py:50) and
// q^1.5 with q in [0,1] (focal_loss.py:44).  x^y = exp(y*log(x)) in fp32; relative error grows
// with |y*log x| (about 1 + |y ln x| ulp), measured in oracle/DETMATH_REPORT.md.
DM_HD float dm_powf(float x, float y) {
  if (dm_isnan(x) || dm_isnan(y)) return dm_u2f(DM_NAN_BITS);
  if (y == 0.0f) return 1.0f;
  if (x == 0.0f) return (y > 0.0f) ? 0.0f : dm_u2f(DM_INF_BITS);
  if (x < 0.0f) return dm_u2f(DM_NAN_BITS);
  if (x == 1.0f) return 1.0f;
  return dm_expf(DM_MUL(y, dm_logf(x)));
}

// q^1.5 = q*sqrt(q) (both correctly rounded steps); used for the focal modulating factor.
DM_HD float dm_pow15f(float q) { return DM_MUL(q, DM_SQRT(q)); }

// BCE-with-logits, op order of tf.nn.sigmoid_cross_entropy_with_logits:
//   (max(x,0) - x*z) + log1p(exp(-|x|))
DM_HD float dm_bce_logits(float z, float x) {
  float relu = (x >= 0.0f) ? x : 0.0f;
  float nabs = (x >= 0.0f) ? -x : x;
  return DM_ADD(DM_SUB(relu, DM_MUL(x, z)), dm_log1pf(dm_expf(nabs)));
}
