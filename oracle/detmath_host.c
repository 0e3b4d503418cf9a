/* oracle/detmath_host.c — TEST INFRASTRUCTURE (oracle side).
 *
 * Host build of the deterministic math header the CUDA kernels use, exposed as array
 * functions so the NumPy oracle can evaluate exp / sigmoid / log / log1p / atan / pow with
 * exactly the bits the device produces.  Nothing in the product path loads this library.
 * Build: `make -C oracle` (gcc -O2 -ffp-contract=off).  */
#include <stddef.h>
#include "../tensorflow2-machine-vision_b200/csrc/detmath.h"

#define VEC1(name, fn)                                        \
  void name(const float* in, float* out, size_t n) {          \
    for (size_t i = 0; i < n; ++i) out[i] = fn(in[i]);        \
  }

VEC1(om_exp, dm_expf)
VEC1(om_sigmoid, dm_sigmoidf)
VEC1(om_log, dm_logf)
VEC1(om_log1p, dm_log1pf)
VEC1(om_atan, dm_atanf)
VEC1(om_pow15, dm_pow15f)

void om_pow(const float* x, float y, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = dm_powf(x[i], y);
}

/* (max(x,0) - x*z) + log1p(exp(-|x|)) elementwise */
void om_bce_logits(const float* z, const float* x, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = dm_bce_logits(z[i], x[i]);
}

int om_abi_version(void) { return 1; }
