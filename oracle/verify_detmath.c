/* oracle/verify_detmath.c — TEST INFRASTRUCTURE.
 *
 * Exhaustive sweep of csrc/detmath.h over every fp32 bit pattern of each function's domain,
 * against double-precision libm rounded to fp32.  Reports max error in ulp, the worst input,
 * and (for sigmoid / exp) whether the function is monotone non-decreasing over ordered floats.
 * Run: make -C oracle verify   (about a minute on 8 cores).  Output pasted into
 * oracle/DETMATH_REPORT.md.  */
#include <stdio.h>
#include <stdlib.h>
#include <float.h>
#include <omp.h>
#include "../tensorflow2-machine-vision_b200/csrc/detmath.h"

static double ulp_of(double ref) {
  /* spacing of fp32 at |ref| (denormal spacing below FLT_MIN) */
  double a = fabs(ref);
  if (a < (double)FLT_MIN) return ldexp(1.0, -149);
  int e;
  frexp(a, &e); /* a = m*2^e, m in [0.5,1) */
  return ldexp(1.0, e - 24);
}

typedef float (*f1)(float);
typedef double (*d1)(double);

static double d_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
static double d_pow06(double x) { return pow(x, 0.6); }
static float f_pow06(float x) { return dm_powf(x, 0.6f); }
static double d_pow15(double x) { return pow(x, 1.5); }
static double d_softplus_neg(double x) { return log1p(exp(-fabs(x))); }
static float f_softplus_neg(float x) { return dm_log1pf(dm_expf(-dm_fabsf(x))); }

static void sweep(const char* name, f1 f, d1 d, float lo, float hi, int check_mono) {
  /* iterate ordered floats from lo to hi via the monotone integer mapping */
  uint32_t ulo = dm_f2u(lo), uhi = dm_f2u(hi);
  int64_t klo = (ulo & 0x80000000u) ? -(int64_t)(ulo & 0x7fffffffu) : (int64_t)ulo;
  int64_t khi = (uhi & 0x80000000u) ? -(int64_t)(uhi & 0x7fffffffu) : (int64_t)uhi;
  double worst = 0.0;
  float worst_x = 0.0f;
  long long mono_viol = 0;
  long long count = khi - klo + 1;
#pragma omp parallel
  {
    double w = 0.0;
    float wx = 0.0f;
    long long mv = 0;
#pragma omp for schedule(static)
    for (int64_t k = klo; k <= khi; ++k) {
      uint32_t u = (k < 0) ? (0x80000000u | (uint32_t)(-k)) : (uint32_t)k;
      float x = dm_u2f(u);
      float y = f(x);
      double ref = d((double)x);
      double err;
      if (isinf(ref) || ref > (double)FLT_MAX) {
        err = isinf(y) ? 0.0 : 1e9;
      } else {
        err = fabs((double)y - ref) / ulp_of(ref);
      }
      if (err > w) { w = err; wx = x; }
      if (check_mono && k < khi) {
        int64_t k2 = k + 1;
        uint32_t u2 = (k2 < 0) ? (0x80000000u | (uint32_t)(-k2)) : (uint32_t)k2;
        float y2 = f(dm_u2f(u2));
        if (y2 < y) mv++;
      }
    }
#pragma omp critical
    {
      if (w > worst) { worst = w; worst_x = wx; }
      mono_viol += mv;
    }
  }
  printf("%-14s inputs=%lld  [%.9g, %.9g]  max_err=%.3f ulp at x=%.9g (0x%08x)", name, count, lo, hi, worst,
         worst_x, dm_f2u(worst_x));
  if (check_mono) printf("  monotone_violations=%lld", mono_viol);
  printf("\n");
  fflush(stdout);
}

int main(void) {
  printf("threads=%d\n", omp_get_max_threads());
  sweep("exp", dm_expf, exp, -104.0f, 88.72283935546875f, 1);
  sweep("sigmoid", dm_sigmoidf, d_sigmoid, -104.0f, 104.0f, 1);
  sweep("log", dm_logf, log, 1.401298464e-45f, FLT_MAX, 1);
  sweep("atan", dm_atanf, atan, -FLT_MAX, FLT_MAX, 1);
  sweep("pow(x,0.6)", f_pow06, d_pow06, 1.401298464e-45f, 1.0f, 1);
  sweep("pow(x,0.6)>1e-6", f_pow06, d_pow06, 1e-6f, 1.0f, 0);
  sweep("x^1.5", dm_pow15f, d_pow15, 1e-30f, 1.0f, 1);
  sweep("log1p(e^-|x|)", f_softplus_neg, d_softplus_neg, -100.0f, 100.0f, 0);
  /* plateau constants used by the decode kernel's class-argmax shortcut */
  {
    float x = 10.0f;
    while (dm_sigmoidf(x) < 1.0f) x = nextafterf(x, 100.0f);
    printf("smallest x with sigmoid(x)==1.0f: %.9g (0x%08x)\n", x, dm_f2u(x));
  }
  return 0;
}
