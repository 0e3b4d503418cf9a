// api_common.cu — error slot, version, device query, and the detmath self-test kernel of libb200det.so.
#include "common.cuh"
#include "detmath.h"
#include "../../include/b200det.h"

static thread_local char g_err[512] = "";

void b200_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int b200_sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

extern "C" const char* b200_last_error(void) { return g_err; }
extern "C" int b200_version(void) { return B200DET_VERSION; }

extern "C" int b200_device_ok(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    b200_set_error("no CUDA device visible (%s); libb200det has no CPU fallback", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

// ---- detmath on the device, for the host==device bit-equality test -------------------------
__global__ void detmath_eval_kernel(int op, const float* __restrict__ x, const float* __restrict__ y,
                                    float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float a = x[i];
    float r;
    switch (op) {
      case B200_DM_EXP: r = dm_expf(a); break;
      case B200_DM_SIGMOID: r = dm_sigmoidf(a); break;
      case B200_DM_LOG: r = dm_logf(a); break;
      case B200_DM_LOG1P: r = dm_log1pf(a); break;
      case B200_DM_ATAN: r = dm_atanf(a); break;
      case B200_DM_POW06: r = dm_powf(a, 0.6f); break;
      case B200_DM_POW15: r = dm_pow15f(a); break;
      case B200_DM_BCE: r = dm_bce_logits(y[i], a); break;
      default: r = 0.0f;
    }
    out[i] = r;
  }
}

extern "C" int b200_detmath_eval(int op, const float* x, const float* y, float* out, size_t n, void* stream) {
  B200_REQUIRE(op >= 0 && op < B200_DM_COUNT, B200_ERR_BAD_ARG, "b200_detmath_eval: bad op %d", op);
  B200_REQUIRE(x && out, B200_ERR_BAD_ARG, "b200_detmath_eval: null pointer");
  B200_REQUIRE(op != B200_DM_BCE || y, B200_ERR_BAD_ARG, "b200_detmath_eval: BCE needs y");
  if (n == 0) return B200_OK;
  int blocks = (int)((n + 255) / 256);
  int cap = b200_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  detmath_eval_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(op, x, y, out, n);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// Device-wide hint: DRAM->L2 fetch granularity (32/64/128 bytes).  The loss kernels touch one 32-byte sector per
// 340-byte record; with the default granularity every touched 128-byte line is fetched whole.
extern "C" int b200_set_l2_fetch_granularity(int bytes) {
  B200_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, B200_ERR_BAD_ARG, "b200_set_l2_fetch_granularity: %d not in {32,64,128}", bytes);
  B200_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return B200_OK;
}
extern "C" int b200_get_l2_fetch_granularity(void) {
  size_t v = 0;
  if (cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) != cudaSuccess) return -1;
  return (int)v;
}
