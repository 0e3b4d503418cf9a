// effdet_focal.cuh — element math of the EfficientDet losses (focal_loss.py:36-52, box_loss.py:17-29), shared by the
// stand-alone loss kernels (effdet_loss.cu) and the fused eval-step stream (effdet.cu).
#pragma once

// single-instruction MUFU approximations (relative error ~1e-7..1e-6): 4 per element, the SFU pipe (16 lanes/SM/clk)
// stays below the HBM time of the two class tensors
__device__ __forceinline__ float el_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float el_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float el_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float el_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// The library is built with -fmad=false (decision paths must round every operation separately), so the fused multiply-adds
// of this continuous-valued loss are written out: 21 instructions per element instead of 28, 4 of them MUFU.
template <bool G15 = false>
__device__ __forceinline__ float el_focal(float y, float x, float alpha, float gamma, float ls) {
  // focal_loss.py:36-52
  const float e = el_ex2(-1.4426950408889634f * fabsf(x));   // exp(-|x|)
  const float r = el_rcp(1.0f + e);
  const float p = (x >= 0.0f) ? r : e * r;                    // sigmoid(x)
  // 1 - p_t with p_t = y p + (1 - y)(1 - p):  p + y (1 - 2p)
  const float q = fmaxf(__fmaf_rn(y, __fmaf_rn(-2.0f, p, 1.0f), p), 0.0f);
  const float af = __fmaf_rn(y, 2.0f * alpha - 1.0f, 1.0f - alpha);   // y alpha + (1 - y)(1 - alpha); the constants fold per kernel
  const float mod = (G15 || gamma == 1.5f) ? q * el_sqrt(q) : __powf(q, gamma);
  const float ys = __fmaf_rn(y, 1.0f - ls, 0.5f * ls);
  // sigmoid_cross_entropy_with_logits: max(x,0) - x ys + log1p(exp(-|x|)), and log1p(e) = -log(1 / (1 + e))
  const float ce = __fmaf_rn(-0.6931471805599453f, el_lg2(r), __fmaf_rn(-x, ys, fmaxf(x, 0.0f)));
  return af * mod * ce;
}

// y == 0 specialisation for the sparse-target mode (every element is background except one per anchor):
// p_t = 1 - p, alpha factor = 1 - alpha, modulating factor = p^gamma, and log1p(exp(-|x|)) as e*P(e) with a degree-6
// polynomial on e in (0,1] (max relative error 1.5e-6) so that only 3 of the 4 MUFU operations remain — the focal
// pass is bound by the SFU pipe (4 lanes per scheduler), not by FP32 issue.
__device__ __forceinline__ float el_log1p_poly(float e) {
  float p = 0.014202825725078583f;
  p = __fmaf_rn(p, e, -0.06658805161714554f);
  p = __fmaf_rn(p, e, 0.14943458139896393f);
  p = __fmaf_rn(p, e, -0.23514863848686218f);
  p = __fmaf_rn(p, e, 0.3311205208301544f);
  p = __fmaf_rn(p, e, -0.4998719096183777f);
  p = __fmaf_rn(p, e, 0.9999987483024597f);
  return p * e;
}

template <bool G15>
__device__ __forceinline__ float el_focal_bg(float x, float one_minus_alpha, float gamma, float half_ls) {
  const float e = el_ex2(-1.4426950408889634f * fabsf(x));   // exp(-|x|)
  const float r = el_rcp(1.0f + e);
  const float p = (x >= 0.0f) ? r : e * r;                   // sigmoid(x) = 1 - p_t
  const float mod = G15 ? p * el_sqrt(p) : __powf(p, gamma);
  const float ce = __fmaf_rn(-x, half_ls, fmaxf(x, 0.0f)) + el_log1p_poly(e);
  return one_minus_alpha * mod * ce;
}

__device__ __forceinline__ float el_huber(float t, float o, float delta) {
  // keras Huber on the size-1 last axis: |e| <= d ? 0.5 e^2 : d|e| - 0.5 d^2, masked by target != 0 (box_loss.py:24)
  if (t == 0.0f) return 0.0f;
  const float e = o - t, a = fabsf(e);
  return (a <= delta) ? 0.5f * e * e : delta * a - 0.5f * delta * delta;
}

