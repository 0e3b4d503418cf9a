// iou.cu — pairwise / elementwise box metrics behind GetIOU (utils/tf_iou_utils.py:5-65) and
// get_iou (efficientnet/utils/iou.py:26-100).
#include "boxmath.cuh"
#include "common.cuh"
#include "../../include/b200det.h"

// out[i*n2 + j] = metric(b1[i], b2[j]); one thread per pair, b2 tile staged through shared memory.
__global__ void pairwise_iou_kernel(const float* __restrict__ b1, int n1, const float* __restrict__ b2, int n2,
                                    int metric, float* __restrict__ out) {
  __shared__ BoxT s2[128];
  const int j0 = blockIdx.x * 128;
  const int i0 = blockIdx.y * 128;
  if (j0 + threadIdx.x < n2) {
    const float* p = b2 + 4 * (size_t)(j0 + threadIdx.x);
    s2[threadIdx.x] = bm_prep(p[0], p[1], p[2], p[3], metric);
  }
  __syncthreads();
  const int jn = min(128, n2 - j0);
  // thread t handles column t of rows i0..i0+127 so that stores along j are coalesced
  if ((int)threadIdx.x < jn) {
    BoxT bj = s2[threadIdx.x];
    const int in = min(128, n1 - i0);
    for (int r = 0; r < in; ++r) {
      const float* p = b1 + 4 * (size_t)(i0 + r);
      BoxT bi = bm_prep(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3), metric);
      out[(size_t)(i0 + r) * n2 + j0 + threadIdx.x] = bm_metric(bi, bj, metric);
    }
  }
}

__global__ void elementwise_iou_kernel(const float* __restrict__ b1, const float* __restrict__ b2, size_t n,
                                       int metric, float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float* p = b1 + 4 * i;
    const float* q = b2 + 4 * i;
    BoxT a = bm_prep(p[0], p[1], p[2], p[3], metric);
    BoxT b = bm_prep(q[0], q[1], q[2], q[3], metric);
    out[i] = bm_metric(a, b, metric);
  }
}

extern "C" int b200_pairwise_iou(const float* b1, int n1, const float* b2, int n2, int metric, float* out,
                                 void* stream) {
  B200_REQUIRE(metric >= 0 && metric < B200_METRIC_COUNT, B200_ERR_BAD_ARG, "b200_pairwise_iou: bad metric %d", metric);
  B200_REQUIRE(n1 >= 0 && n2 >= 0, B200_ERR_BAD_ARG, "b200_pairwise_iou: negative size");
  if (n1 == 0 || n2 == 0) return B200_OK;
  B200_REQUIRE(b1 && b2 && out, B200_ERR_BAD_ARG, "b200_pairwise_iou: null pointer");
  dim3 grid((n2 + 127) / 128, (n1 + 127) / 128);
  B200_REQUIRE(grid.y <= 65535, B200_ERR_BAD_ARG, "b200_pairwise_iou: n1 too large (%d)", n1);
  pairwise_iou_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(b1, n1, b2, n2, metric, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_elementwise_iou(const float* b1, const float* b2, size_t n, int metric, float* out,
                                    void* stream) {
  B200_REQUIRE(metric >= 0 && metric < B200_METRIC_COUNT, B200_ERR_BAD_ARG, "b200_elementwise_iou: bad metric %d", metric);
  if (n == 0) return B200_OK;
  B200_REQUIRE(b1 && b2 && out, B200_ERR_BAD_ARG, "b200_elementwise_iou: null pointer");
  size_t blocks = (n + 255) / 256;
  size_t cap = (size_t)b200_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  elementwise_iou_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(b1, b2, n, metric, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// _get_v of the EfficientDet CIoU (efficientnet/utils/iou.py:5-24) with its tf.custom_gradient: v = 4 ((atan(dnn(w1, h1)) -
// atan(dnn(w2, h2))) / pi)^2 and, for an upstream gradient dv, the reference's hand-written gradient with respect to the
// second box's (height, width): gdh = -dv * 8 * arctan * w2 / pi^2, gdw = dv * 8 * arctan * h2 / pi^2 (:17-18; the 1 / (w^2 +
// h^2) factor of the true derivative is deliberately absent there, as in google/automl).  Elementwise over n boxes; every
// step one rounded fp32 operation in the reference's order.
__global__ void ciou_v_grad_kernel(const float* __restrict__ h1, const float* __restrict__ w1, const float* __restrict__ h2,
                                   const float* __restrict__ w2, const float* __restrict__ dv, size_t n, float* __restrict__ v,
                                   float* __restrict__ gdh, float* __restrict__ gdw) {
  const float pi = B200_PI_F;   // math.pi**2 below is a Python double (9.869604401089358) rounded to fp32 when it meets the tensor
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float arct = DM_SUB(dm_atanf(bm_dnn(w1[i], h1[i])), dm_atanf(bm_dnn(w2[i], h2[i])));
    const float q = DM_DIV(arct, pi);
    if (v) v[i] = DM_MUL(4.0f, DM_MUL(q, q));
    if (dv) {
      const float d = dv[i];
      // evaluation order of :17-18: ((dv * 8) * arctan) * height / pi^2 ; ((-dv * 8) * arctan) * width / pi^2
      if (gdw) gdw[i] = DM_DIV(DM_MUL(DM_MUL(DM_MUL(d, 8.0f), arct), h2[i]), 9.869604401089358f);
      if (gdh) gdh[i] = DM_DIV(DM_MUL(DM_MUL(DM_MUL(-d, 8.0f), arct), w2[i]), 9.869604401089358f);
    }
  }
}

extern "C" int b200_ciou_v_grad(const float* b1_height, const float* b1_width, const float* b2_height, const float* b2_width,
                                const float* dv, size_t n, float* v_out, float* grad_height_out, float* grad_width_out,
                                void* stream) {
  if (n == 0) return B200_OK;
  B200_REQUIRE(b1_height && b1_width && b2_height && b2_width, B200_ERR_BAD_ARG, "b200_ciou_v_grad: null pointer");
  B200_REQUIRE(v_out || (dv && (grad_height_out || grad_width_out)), B200_ERR_BAD_ARG, "b200_ciou_v_grad: nothing to compute");
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)b200_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ciou_v_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(b1_height, b1_width, b2_height, b2_width, dv, n, v_out,
                                                                    grad_height_out, grad_width_out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
