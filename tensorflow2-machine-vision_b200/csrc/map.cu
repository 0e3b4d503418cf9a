// map.cu — per-image mAP, the step right after NMS in test_step.
//
// Replaces utils/mAP.py:3-125 (Get_TPFP / Get_AP / Get_mAP_one), which the reference reaches through
// tf.numpy_function (a GIL round trip with a Python loop over the classes; yolo_v4/model.py:377, edt:166).
// One CTA per image, fp64 like the NumPy original.  Per class: compact the class's predictions and ground truth
// (order preserved), one thread per ground-truth box finds its arg-max-IoU prediction (first maximum) and marks it
// a true positive when IoU >= thresh, a rank sort orders the predictions by score (descending; equal scores in
// reversed index order, as a stable argsort followed by [::-1] gives), and one thread walks the list to build the
// reference's precision/recall arrays, their envelope and the AP sum exactly as written (including its swapped
// mrec/mpre naming, mAP.py:88-99).
#include "common.cuh"

#define MAP_THREADS 256

struct MapParams {
  const float* gt; const int32_t* gt_off;   // [total_gt,5] x1,y1,x2,y2,class ; [B+1]
  const float* pr; const int32_t* pr_off;   // [total_pr,6] x1,y1,x2,y2,class,score ; [B+1]
  int class_num; double thresh;
  int cap_p, cap_g;                         // per-image capacities of the shared lists
  double* out;                              // [B]
};

__global__ void __launch_bounds__(MAP_THREADS) map_kernel(MapParams p) {
  extern __shared__ __align__(16) unsigned char map_smem[];
  double* s_a = reinterpret_cast<double*>(map_smem);        // [cap_p + 2] "mrec" (precision list)
  double* s_b = s_a + p.cap_p + 2;                          // [cap_p + 2] "mpre" (recall list)
  int* s_pi = reinterpret_cast<int*>(s_b + p.cap_p + 2);    // [cap_p] prediction rows of the class
  int* s_gi = s_pi + p.cap_p;                               // [cap_g] ground-truth rows of the class
  int* s_tp = s_gi + p.cap_g;                               // [cap_p] tp flag per class prediction
  int* s_sorted = s_tp + p.cap_p;                           // [cap_p] tp flags in sorted order
  __shared__ int s_np, s_ng;
  __shared__ double s_total;
  __shared__ int s_warp[MAP_THREADS / 32];
  const int img = blockIdx.x;
  const int g0 = p.gt_off[img], ng_all = p.gt_off[img + 1] - g0;
  const int p0 = p.pr_off[img], np_all = p.pr_off[img + 1] - p0;
  const float* gt = p.gt + 5 * (size_t)g0;
  const float* pr = p.pr + 6 * (size_t)p0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_total = 0.0;
  __syncthreads();
  for (int c = 0; c < p.class_num; ++c) {
    // ---- ordered compaction of the class's predictions and ground truth ----
    for (int which = 0; which < 2; ++which) {
      const int n_all = which ? ng_all : np_all;
      const float* base = which ? gt : pr;
      const int stride = which ? 5 : 6;
      int* dst = which ? s_gi : s_pi;
      int run = 0;
      for (int i0 = 0; i0 < n_all; i0 += MAP_THREADS) {
        const int i = i0 + tid;
        const bool f = (i < n_all) && ((double)base[(size_t)i * stride + 4] == (double)c);
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = run;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (f) dst[off + __popc(bal & ((1u << lane) - 1u))] = i;
        int tot = 0;
        for (int w = 0; w < MAP_THREADS / 32; ++w) tot += s_warp[w];
        run += tot;
        __syncthreads();
      }
      if (tid == 0) { if (which) s_ng = run; else s_np = run; }
    }
    __syncthreads();
    const int np_c = s_np, ng_c = s_ng;
    if (np_c == 0 || ng_c == 0) { __syncthreads(); continue; }  // AP = 0 (mAP.py:29-30 -> empty tp list)
    for (int i = tid; i < np_c; i += MAP_THREADS) s_tp[i] = 0;
    __syncthreads();
    // ---- each ground-truth box claims its arg-max-IoU prediction ----
    for (int g = tid; g < ng_c; g += MAP_THREADS) {
      const float* gb = gt + 5 * (size_t)s_gi[g];
      const double gx1 = gb[0], gy1 = gb[1], gx2 = gb[2], gy2 = gb[3];
      const double ga = (gx2 - gx1) * (gy2 - gy1);
      int best = 0;
      double bv = 0.0;
      for (int q = 0; q < np_c; ++q) {
        const float* pb = pr + 6 * (size_t)s_pi[q];
        const double px1 = pb[0], py1 = pb[1], px2 = pb[2], py2 = pb[3];
        const double iw = fmax(fmin(gx2, px2) - fmax(gx1, px1), 0.0), ih = fmax(fmin(gy2, py2) - fmax(gy1, py1), 0.0);
        const double inter = iw * ih;
        const double iou = inter / (ga + (px2 - px1) * (py2 - py1) - inter);
        // np.argmax (mAP.py:53): the first maximal index, and a NaN (0/0: zero-area boxes) counts as maximal — the
        // first NaN freezes the choice, and `NaN >= thresh` is False, so that ground-truth box claims nothing
        if (q == 0 || (bv == bv && (iou > bv || iou != iou))) { bv = iou; best = q; }
      }
      if (bv >= p.thresh) s_tp[best] = 1;
    }
    __syncthreads();
    // ---- rank sort by score, descending; ties: higher original position first ----
    for (int i = tid; i < np_c; i += MAP_THREADS) {
      const float si = pr[6 * (size_t)s_pi[i] + 5];
      int rank = 0;
      for (int j = 0; j < np_c; ++j) {
        const float sj = pr[6 * (size_t)s_pi[j] + 5];
        rank += (sj > si || (sj == si && j > i)) ? 1 : 0;
      }
      s_sorted[rank] = s_tp[i];
    }
    __syncthreads();
    // ---- AP exactly as mAP.py:74-99 ----
    if (tid == 0) {
      double sum = 0.0;
      s_a[0] = 0.0; s_b[0] = 0.0;
      for (int i = 0; i < np_c; ++i) {
        if (s_sorted[i] == 1) sum += 1.0;
        s_a[i + 1] = sum / (double)(i + 1);   // "precision_list" -> mrec
        s_b[i + 1] = sum / (double)ng_c;      // "recall_list"    -> mpre
      }
      s_a[np_c + 1] = 1.0; s_b[np_c + 1] = 0.0;
      for (int i = np_c + 1; i > 0; --i) s_b[i - 1] = fmax(s_b[i - 1], s_b[i]);
      double ap = 0.0;
      for (int i = 0; i <= np_c; ++i)
        if (s_a[i + 1] != s_a[i]) ap += (s_a[i + 1] - s_a[i]) * s_b[i + 1];
      s_total += ap;
    }
    __syncthreads();
  }
  if (tid == 0) p.out[img] = s_total / (double)p.class_num;
}

extern "C" int b200_map_per_image(const float* gt, const int32_t* gt_offsets, const float* pred,
                                  const int32_t* pred_offsets, int num_images, int max_gt_per_image,
                                  int max_pred_per_image, int class_num, double thresh, double* out, void* stream) {
  B200_REQUIRE(num_images >= 0 && class_num >= 1 && max_gt_per_image >= 0 && max_pred_per_image >= 0, B200_ERR_BAD_ARG,
               "b200_map_per_image: bad sizes");
  if (num_images == 0) return B200_OK;
  B200_REQUIRE(gt_offsets && pred_offsets && out, B200_ERR_BAD_ARG, "b200_map_per_image: null pointer");
  MapParams p;
  p.gt = gt; p.gt_off = gt_offsets; p.pr = pred; p.pr_off = pred_offsets; p.class_num = class_num; p.thresh = thresh;
  p.cap_p = max_pred_per_image > 0 ? max_pred_per_image : 1;
  p.cap_g = max_gt_per_image > 0 ? max_gt_per_image : 1;
  p.out = out;
  const size_t smem = sizeof(double) * 2 * (size_t)(p.cap_p + 2) + sizeof(int) * (3 * (size_t)p.cap_p + p.cap_g);
  B200_REQUIRE(smem <= 200 * 1024, B200_ERR_UNSUPPORTED, "b200_map_per_image: too many boxes per image (%d predictions, %d ground truth)",
               max_pred_per_image, max_gt_per_image);
  B200_CUDA(cudaFuncSetAttribute(map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  map_kernel<<<num_images, MAP_THREADS, smem, (cudaStream_t)stream>>>(p);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
