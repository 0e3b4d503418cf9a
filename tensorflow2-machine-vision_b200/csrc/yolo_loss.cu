// yolo_loss.cu — fused YOLOv3/v4 loss: GetLoss (utils/tf_yolo_utils.py:6-127) and its keras-yolo3 twin
// Yolov4Loss.call (losses/yolo_loss.py:85-159).
//
// The reference materialises ~20 (H,W,A,n_gt) broadcast temporaries per image inside a tf.while_loop; here
// the loss is three launches over the dense NHWC tensors, touching only the sectors the arithmetic needs:
//
//  K4a yolo_loss_objects_kernel   one CTA per 256 consecutive anchor records of one (level, image):
//      reads the obj channel of y_true (one 32-byte sector per 340-byte record), stores it to a compact
//      per-anchor array, and for records with obj != 0 (the ground-truth list, tyu:82) a warp per record reads
//      the full y_true / y_pred records and produces the xy, wh and class terms (tyu:107-118) and the
//      prepared GT box (corners, area, atan(w/h)) appended to the (image, level) GT list.
//  K4b yolo_loss_ignore_kernel    same grid: reads the 5 box/conf logits of every y_pred record, decodes the
//      predicted box (tyu:57-75), tests it against the GT list held in shared memory with the cheap
//      "no overlap => metric <= 0 < thr" reject before any exact metric (iou / diou / ciou, tiu:5-65), and
//      accumulates object_loss = obj*bce + (1-obj)*bce*ignore (tyu:111-114).
//  K4c yolo_loss_finalize_kernel  fixed-order fp64 reduction of the per-CTA partials -> parts[3][4] / batch,
//      loss = sum_l ((xy+wh)+obj)+cls in fp32 in the reference's order (tyu:120-125).  Deterministic run to run.
//
// ignore = float(best < thr) with best = max_g metric(pred, gt_g) is evaluated as "no g with metric >= thr"
// (identical unless a metric is NaN, which needs non-finite boxes).
#include "boxmath.cuh"
#include "common.cuh"

#define YL_LEVELS 3
#define YL_CHUNK 256

enum { YL_VARIANT_TF_YOLO_UTILS = 0, YL_VARIANT_KERAS_YOLO3 = 1 };

struct YlLevels {
  const float* y_true[YL_LEVELS];
  const float* y_pred[YL_LEVELS];
  int h[YL_LEVELS], w[YL_LEVELS], rec_per_img[YL_LEVELS], anchor_base[YL_LEVELS];
  int chunks_per_img[YL_LEVELS];
  int cta_base[YL_LEVELS + 1];
  float anc_w[YL_LEVELS][8], anc_h[YL_LEVELS][8];  // pixels
};

struct YlParams {
  YlLevels lv;
  int B, A, C, RF, n_img;
  float img_w, img_h;
  float thr;
  int metric, variant;
  float* obj_compact;   // [B, n_img]
  BoxT* gt;             // [B, n_img] (level slices at anchor_base)
  int32_t* gt_count;    // [B, 3]
  double* partials;     // [n_cta, 4]  xy, wh, obj, cls
};

__device__ __forceinline__ void yl_locate(const YlLevels& lv, int cta, int& l, int& img, int& chunk) {
  l = 0;
#pragma unroll
  for (int k = 1; k < YL_LEVELS; ++k) if (cta >= lv.cta_base[k]) l = k;
  const int r = cta - lv.cta_base[l];
  img = r / lv.chunks_per_img[l];
  chunk = r - img * lv.chunks_per_img[l];
}

__global__ void __launch_bounds__(YL_CHUNK) yolo_loss_objects_kernel(YlParams p) {
  __shared__ int s_list[YL_CHUNK];
  __shared__ int s_n;
  __shared__ double s_acc[YL_CHUNK / 32][3];
  int l, img, chunk;
  yl_locate(p.lv, blockIdx.x, l, img, chunk);
  const int rpi = p.lv.rec_per_img[l];
  const int rin = chunk * YL_CHUNK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const float* yt = p.lv.y_true[l] + ((size_t)img * rpi) * p.RF;
  const float* yp = p.lv.y_pred[l] + ((size_t)img * rpi) * p.RF;
  if (rin < rpi) {
    const float obj = __ldg(yt + (size_t)rin * p.RF + 4);
    p.obj_compact[(size_t)img * p.n_img + p.lv.anchor_base[l] + rin] = obj;
    if (obj != 0.0f) s_list[atomicAdd(&s_n, 1)] = rin;
  }
  __syncthreads();
  const int n = s_n;
  double a_xy = 0.0, a_wh = 0.0, a_cls = 0.0;
  const int W = p.lv.w[l], H = p.lv.h[l];
  for (int k = warp; k < n; k += YL_CHUNK / 32) {
    const int r = s_list[k];
    const float* t = yt + (size_t)r * p.RF;
    const float* q = yp + (size_t)r * p.RF;
    const float obj = __ldg(t + 4);
    const float tx = __ldg(t), ty = __ldg(t + 1), tw = __ldg(t + 2), th = __ldg(t + 3);
    const float scale = DM_SUB(2.0f, DM_MUL(tw, th));
    const float os = DM_MUL(obj, scale);  // (obj * scale) * term, evaluation order of tyu:109-110
    float e_xy = 0.f, e_wh = 0.f, e_cls = 0.f;
    if (lane < 2) {
      const int cell = r / p.A;
      const int gy = cell / W, gx = cell - gy * W;
      const float tv = lane == 0 ? tx : ty;
      const float g = lane == 0 ? (float)gx : (float)gy;
      const float gs = lane == 0 ? (float)W : (float)H;
      float raw = DM_SUB(DM_MUL(tv, gs), g);
      if (p.variant == YL_VARIANT_TF_YOLO_UTILS) raw = DM_MUL(obj, raw);  // tyu:44 (absent in yolo_loss.py:120)
      e_xy = DM_MUL(os, dm_bce_logits(raw, __ldg(q + lane)));
    } else if (lane < 4) {
      const int a = r - (r / p.A) * p.A;
      const float tv = lane == 2 ? tw : th;
      const float im = lane == 2 ? p.img_w : p.img_h;
      const float an = lane == 2 ? p.lv.anc_w[l][a] : p.lv.anc_h[l][a];
      float num = DM_MUL(tv, im);
      if (p.variant == YL_VARIANT_TF_YOLO_UTILS) num = DM_ADD(num, 1e-8f);  // tyu:48
      const float raw = dm_logf(DM_DIV(num, an));
      const float d = DM_SUB(raw, __ldg(q + lane));
      e_wh = DM_MUL(DM_MUL(os, 0.5f), DM_MUL(d, d));
    }
    for (int c = 5 + lane; c < p.RF; c += 32) e_cls += DM_MUL(obj, dm_bce_logits(__ldg(t + c), __ldg(q + c)));
    e_xy = warp_sum(e_xy); e_wh = warp_sum(e_wh); e_cls = warp_sum(e_cls);
    if (lane == 0) {
      a_xy += (double)e_xy; a_wh += (double)e_wh; a_cls += (double)e_cls;
      // ground-truth box for the ignore mask: corners of (t_xy, t_wh), tyu:68-71
      const float hx = DM_DIV(tw, 2.0f), hy = DM_DIV(th, 2.0f);
      BoxT g = bm_prep(DM_SUB(tx, hx), DM_SUB(ty, hy), DM_ADD(tx, hx), DM_ADD(ty, hy), p.metric);
      const int slot = atomicAdd(&p.gt_count[img * YL_LEVELS + l], 1);
      p.gt[(size_t)img * p.n_img + p.lv.anchor_base[l] + slot] = g;
    }
  }
  if (lane == 0) { s_acc[warp][0] = a_xy; s_acc[warp][1] = a_wh; s_acc[warp][2] = a_cls; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0, w = 0, c = 0;
    for (int i = 0; i < YL_CHUNK / 32; ++i) { x += s_acc[i][0]; w += s_acc[i][1]; c += s_acc[i][2]; }
    double* out = p.partials + (size_t)blockIdx.x * 4;
    out[0] = x; out[1] = w; out[3] = c;
  }
}

#define YL_GT_TILE 128

__global__ void __launch_bounds__(YL_CHUNK) yolo_loss_ignore_kernel(YlParams p) {
  __shared__ BoxT s_gt[YL_GT_TILE];
  __shared__ double s_acc[YL_CHUNK / 32];
  int l, img, chunk;
  yl_locate(p.lv, blockIdx.x, l, img, chunk);
  const int rpi = p.lv.rec_per_img[l];
  const int rin = chunk * YL_CHUNK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool active = rin < rpi;
  const int n_gt = p.gt_count[img * YL_LEVELS + l];
  const BoxT* gt = p.gt + (size_t)img * p.n_img + p.lv.anchor_base[l];
  float obj = 0.f, pobj = 0.f;
  BoxT pb;
  pb.c0 = pb.c1 = pb.c2 = pb.c3 = pb.area = pb.at = 0.f;
  if (active) {
    const float* q = p.lv.y_pred[l] + ((size_t)img * rpi + rin) * p.RF;
    obj = p.obj_compact[(size_t)img * p.n_img + p.lv.anchor_base[l] + rin];
    pobj = __ldg(q + 4);
    if (n_gt > 0) {
      const int W = p.lv.w[l], H = p.lv.h[l];
      const int cell = rin / p.A, a = rin - cell * p.A;
      const int gy = cell / W, gx = cell - gy * W;
      // tyu:57,61: xy = (sigmoid(t)+grid)/grid_wh ; wh = exp(t)*anchor/image_wh (no inf guard here, Q7)
      const float x = DM_DIV(DM_ADD(dm_sigmoidf(__ldg(q)), (float)gx), (float)W);
      const float y = DM_DIV(DM_ADD(dm_sigmoidf(__ldg(q + 1)), (float)gy), (float)H);
      const float w = DM_DIV(DM_MUL(dm_expf(__ldg(q + 2)), p.lv.anc_w[l][a]), p.img_w);
      const float h = DM_DIV(DM_MUL(dm_expf(__ldg(q + 3)), p.lv.anc_h[l][a]), p.img_h);
      const float hx = DM_DIV(w, 2.0f), hy = DM_DIV(h, 2.0f);
      pb = bm_prep(DM_SUB(x, hx), DM_SUB(y, hy), DM_ADD(x, hx), DM_ADD(y, hy), p.metric);
    }
  }
  bool hit = false;  // some GT with metric >= thr
  for (int g0 = 0; g0 < n_gt; g0 += YL_GT_TILE) {
    const int m = min(YL_GT_TILE, n_gt - g0);
    __syncthreads();
    for (int i = threadIdx.x; i < m * 6; i += YL_CHUNK)
      reinterpret_cast<float*>(s_gt)[i] = reinterpret_cast<const float*>(gt + g0)[i];
    __syncthreads();
    if (active && !hit) {
      for (int g = 0; g < m; ++g) {
        const BoxT gb = s_gt[g];
        if (bm_surely_below(pb, gb, p.metric, p.thr)) continue;
        if (bm_metric(pb, gb, p.metric) >= p.thr) { hit = true; break; }
      }
    }
  }
  float e = 0.f;
  if (active) {
    const float bc = dm_bce_logits(obj, pobj);
    const float ign = hit ? 0.0f : 1.0f;
    // obj*bc + (1-obj)*bc*ignore, tyu:114
    e = DM_ADD(DM_MUL(obj, bc), DM_MUL(DM_MUL(DM_SUB(1.0f, obj), bc), ign));
  }
  double d = warp_sum_d((double)e);
  if (lane == 0) s_acc[warp] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < YL_CHUNK / 32; ++i) s += s_acc[i];
    p.partials[(size_t)blockIdx.x * 4 + 2] = s;
  }
}

__global__ void __launch_bounds__(1024) yolo_loss_finalize_kernel(const double* __restrict__ partials, int cta_base0,
                                                                   int cta_base1, int cta_base2, int cta_base3,
                                                                   float batch_divisor, float* __restrict__ parts,
                                                                   float* __restrict__ loss) {
  __shared__ double s_red[32][12];
  const int base[4] = {cta_base0, cta_base1, cta_base2, cta_base3};
  double acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.0;
  for (int l = 0; l < 3; ++l)
    for (int c = base[l] + threadIdx.x; c < base[l + 1]; c += blockDim.x) {
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[l * 4 + t] += partials[(size_t)c * 4 + t];
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    double v = warp_sum_d(acc[i]);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.0f;
    for (int l = 0; l < 3; ++l) {
      float t4[4];
      for (int t = 0; t < 4; ++t) {
        double s = 0.0;
        for (int w = 0; w < 32; ++w) s += s_red[w][l * 4 + t];
        t4[t] = DM_DIV((float)s, batch_divisor);  // reduce_sum(...) / batch_size_float, tyu:120-123
        if (parts) parts[l * 4 + t] = t4[t];
      }
      total = DM_ADD(total, DM_ADD(DM_ADD(DM_ADD(t4[0], t4[1]), t4[2]), t4[3]));  // tyu:125
    }
    *loss = total;
  }
}

// ---- host side ---------------------------------------------------------------------------------
struct YlWs { size_t obj, gt, cnt, part, total; int n_cta; };

static YlWs yl_layout(const int32_t hw[6], int B, int A, int* n_img_out, YlLevels* lv) {
  YlWs w;
  int n_img = 0, cta = 0;
  for (int l = 0; l < YL_LEVELS; ++l) {
    const int rpi = hw[2 * l] * hw[2 * l + 1] * A;
    const int cpi = (rpi + YL_CHUNK - 1) / YL_CHUNK;
    if (lv) {
      lv->h[l] = hw[2 * l]; lv->w[l] = hw[2 * l + 1]; lv->rec_per_img[l] = rpi; lv->anchor_base[l] = n_img;
      lv->chunks_per_img[l] = cpi; lv->cta_base[l] = cta;
    }
    n_img += rpi;
    cta += cpi * B;
  }
  if (lv) lv->cta_base[YL_LEVELS] = cta;
  if (n_img_out) *n_img_out = n_img;
  w.n_cta = cta;
  size_t o = 0;
  w.cnt = o; o = b200_align_up(o + sizeof(int32_t) * (size_t)B * YL_LEVELS, 256);
  w.obj = o; o = b200_align_up(o + sizeof(float) * (size_t)B * n_img, 256);
  w.gt = o; o = b200_align_up(o + sizeof(BoxT) * (size_t)B * n_img, 256);
  w.part = o; o = b200_align_up(o + sizeof(double) * 4 * (size_t)cta, 256);
  w.total = o;
  return w;
}

extern "C" size_t b200_yolo_loss_workspace_bytes(const int32_t hw[6], int B, int A) {
  return yl_layout(hw, B, A, nullptr, nullptr).total;
}

extern "C" int b200_yolo_loss(const float* const y_true[3], const float* const y_pred[3], const int32_t hw[6], int B,
                              int A, int C, const float* anchors_wh_host, const float* image_wh_host,
                              float iou_thresh, int metric, int variant, float batch_divisor, float* out_parts,
                              float* out_loss, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_REQUIRE(y_true && y_pred && hw && anchors_wh_host && image_wh_host && out_loss, B200_ERR_BAD_ARG, "b200_yolo_loss: null argument");
  B200_REQUIRE(B >= 1 && A >= 1 && A <= 8 && C >= 0, B200_ERR_BAD_ARG, "b200_yolo_loss: unsupported shape B=%d A=%d C=%d", B, A, C);
  B200_REQUIRE(metric >= B200_METRIC_YOLO_IOU && metric <= B200_METRIC_YOLO_CIOU, B200_ERR_BAD_ARG,
               "b200_yolo_loss: iou_type must be iou/diou/ciou (metric %d)", metric);
  B200_REQUIRE(variant == YL_VARIANT_TF_YOLO_UTILS || variant == YL_VARIANT_KERAS_YOLO3, B200_ERR_BAD_ARG, "b200_yolo_loss: bad variant %d", variant);
  B200_REQUIRE(batch_divisor > 0.0f, B200_ERR_BAD_ARG, "b200_yolo_loss: batch_divisor must be positive");
  YlParams p;
  int n_img = 0;
  YlWs ws = yl_layout(hw, B, A, &n_img, &p.lv);
  B200_REQUIRE(workspace && workspace_bytes >= ws.total, B200_ERR_WORKSPACE, "b200_yolo_loss: workspace %zu < required %zu", workspace_bytes, ws.total);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, B200_ERR_BAD_ARG, "b200_yolo_loss: workspace not 256-byte aligned");
  for (int l = 0; l < YL_LEVELS; ++l) {
    B200_REQUIRE(y_true[l] && y_pred[l] && hw[2 * l] > 0 && hw[2 * l + 1] > 0, B200_ERR_BAD_ARG, "b200_yolo_loss: bad level %d", l);
    p.lv.y_true[l] = y_true[l];
    p.lv.y_pred[l] = y_pred[l];
    for (int a = 0; a < A; ++a) {
      p.lv.anc_w[l][a] = anchors_wh_host[(l * A + a) * 2 + 0];
      p.lv.anc_h[l][a] = anchors_wh_host[(l * A + a) * 2 + 1];
    }
  }
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  p.B = B; p.A = A; p.C = C; p.RF = 5 + C; p.n_img = n_img;
  p.img_w = image_wh_host[0]; p.img_h = image_wh_host[1];
  p.thr = iou_thresh; p.metric = metric; p.variant = variant;
  p.obj_compact = reinterpret_cast<float*>(wsb + ws.obj);
  p.gt = reinterpret_cast<BoxT*>(wsb + ws.gt);
  p.gt_count = reinterpret_cast<int32_t*>(wsb + ws.cnt);
  p.partials = reinterpret_cast<double*>(wsb + ws.part);
  B200_CUDA(cudaMemsetAsync(wsb + ws.cnt, 0, sizeof(int32_t) * (size_t)B * YL_LEVELS, stream));
  B200_CUDA(cudaMemsetAsync(wsb + ws.part, 0, sizeof(double) * 4 * (size_t)ws.n_cta, stream));
  yolo_loss_objects_kernel<<<ws.n_cta, YL_CHUNK, 0, stream>>>(p);
  B200_LAUNCH_CHECK();
  yolo_loss_ignore_kernel<<<ws.n_cta, YL_CHUNK, 0, stream>>>(p);
  B200_LAUNCH_CHECK();
  yolo_loss_finalize_kernel<<<1, 1024, 0, stream>>>(p.partials, p.lv.cta_base[0], p.lv.cta_base[1], p.lv.cta_base[2],
                                                    p.lv.cta_base[3], batch_divisor, out_parts, out_loss);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
