// yolo_decode.cu — YOLOv3/v4 head decode + threshold filter + per-class NMS, fused post-processing.
//
// Replaces GetBoxes (utils/tf_yolo_utils.py:129-167) and GetNMSBoxes (:169-269) of the reference, with the
// B == 1 semantics of the reference applied per image (the reference flattens the batch, tyu:163-166).
//
// Kernel 1  yolo_decode_filter_kernel  (HBM-read bound: every head byte is read exactly once)
//   * each level tensor is treated as a flat array of records of RF = 5+C floats; a warp owns tiles of 32
//     records (32*RF*4 bytes, always a multiple of 16) staged into its own shared-memory ring by 1-D bulk
//     async copies (cp.async.bulk + mbarrier, SASS UBLKCP); 16 warps per SM, each with its own tile in flight;
//   * lane <-> record; the stride RF between lanes is odd for the 85-float COCO record so the column reads are
//     bank-conflict free; conf is thresholded first, then the class maximum is found on raw logits and the
//     sigmoid is evaluated only for logits inside a guard band of the maximum (max / first-argmax are taken
//     in sigmoid space, exactly as tf.reduce_max / tf.argmax of sigmoid(classes) do; detmath's sigmoid is not
//     monotone at ulp level, see oracle/DETMATH_REPORT.md);
//   * survivors (valid box, conf > thr, score > thr: strict, tyu:163,191-192) are appended per image with
//     warp-aggregated atomics.  Append order is arbitrary; each candidate carries its flat anchor index
//     (level-major, then h, w, a) which is monotone in the reference's compaction order, so it serves as the
//     NMS tie-break id and as the rank source for `sel_idx`.
// Kernel 2  yolo_nms_finalize_kernel  (one CTA per image): nms.cuh greedy NMS + gather of the five outputs,
//   recomputing sigmoid(classes) for the <= max_out selected rows straight from the head tensors.
#include <cmath>
#include "nms.cuh"

#define YD_MAX_LEVELS 3
#define YD_STAGES 1

struct YoloLevels {
  const float* head[YD_MAX_LEVELS];
  int h[YD_MAX_LEVELS], w[YD_MAX_LEVELS];
  int rec_per_img[YD_MAX_LEVELS];   // h*w*A
  int anchor_base[YD_MAX_LEVELS];   // flat anchor index of the level's first record within an image
  long long total_rec[YD_MAX_LEVELS];  // B*h*w*A
  long long tile_base[YD_MAX_LEVELS + 1];  // first global tile index of each level
  float anc_w[YD_MAX_LEVELS][8], anc_h[YD_MAX_LEVELS][8];  // anchors_wh / image_wh (fp32 division, tyu:185)
  // lowest tw / th for which the decoded box is certainly valid (w >= 4e-7 so that x + w/2 > x - w/2 in fp32 for any
  // centre in [0,1]); together with t <= 80 (no overflow) and non-NaN tx, ty the filter can skip the decode
  float tmin_w[YD_MAX_LEVELS][8], tmin_h[YD_MAX_LEVELS][8];
};

struct YoloDecodeParams {
  YoloLevels lv;
  int B, A, C, RF;
  int n_img;  // anchors per image over all levels
  float conf_thr, score_thr;
  float conf_lo, conf_hi;  // logits below / above which sigmoid(conf) > conf_thr is decided without the sigmoid
  uint32_t magic_a;        // floor(2^32/A)+1 (0 for A == 1): rin / A == umulhi(rin, magic_a) for rin * A < 2^32
  // candidate store, stride n_img per image
  float4* cand_box; float* cand_score; int32_t* cand_cls; float* cand_conf; uint32_t* cand_aidx;
  int32_t* counts;     // [B]
  uint32_t* bitmap;    // [B, bitmap_words] or nullptr
  int bitmap_words;
  int early_issue;     // refill a warp's slab before its append (atomic + stores) instead of after it
};

// ---- mbarrier / bulk-copy PTX -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// max and first-argmax of sigmoid(x_c), c in [0,C), for the warp's 32 records (lane <-> record, stride RF in shared
// memory, conflict-free).  One pass per lane keeps the largest logit with its first index and the second largest
// value (four independent chains).  When the runner-up is outside the guard band of the maximum its sigmoid is
// certainly smaller and the answer is (sigmoid(m1), i1): for -80 < m < 8, d ln(sigmoid)/dx >= 3.3e-4, so logits
// below m - 0.01 are > 3.3e-6 relative below sigmoid(m) while detmath's sigmoid is within 2.4 ulp = 2.9e-7 of exact
// (oracle/DETMATH_REPORT.md; it is not monotone at ulp level, which is why the band exists).  Otherwise (a few per
// cent of the records) the whole warp evaluates that record's in-band logits and takes max / first-argmax in sigmoid
// space, exactly as tf.reduce_max / tf.argmax of sigmoid(classes) do.
__device__ __forceinline__ void class_max_sigmoid_warp(const float* __restrict__ slab, int RF, int C, bool want, int lane,
                                                       float& best_s, int& best_c) {
  const float* __restrict__ cls = slab + lane * RF + 5;
  bool slow = false;
  float lo = -INFINITY;
  if (want) {
    float a[4], m2[4], nan_acc = 0.0f;
    int ix[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k] = -INFINITY; m2[k] = -INFINITY; ix[k] = C; }
    int c = 0;
    for (; c + 4 <= C; c += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = cls[c + k];
        nan_acc = __fmaf_rn(x, 0.0f, nan_acc);            // NaN once any logit is NaN or +-inf
        m2[k] = fmaxf(m2[k], fminf(a[k], x));
        const bool g = x > a[k];                            // strict: the first index of a chain's maximum stays
        a[k] = g ? x : a[k];
        ix[k] = g ? c + k : ix[k];
      }
    }
    for (; c < C; ++c) {
      const float x = cls[c];
      nan_acc = __fmaf_rn(x, 0.0f, nan_acc);
      m2[0] = fmaxf(m2[0], fminf(a[0], x));
      const bool g = x > a[0];
      a[0] = g ? x : a[0];
      ix[0] = g ? c : ix[0];
    }
    // merge the chains: largest value, lowest index on ties; runner-up = largest of the chains' runner-ups and of
    // the chain maxima that lost
    float m1 = a[0], r2 = m2[0];
    int i1 = ix[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const bool g = (a[k] > m1) || (a[k] == m1 && ix[k] < i1);
      r2 = fmaxf(r2, fmaxf(m2[k], g ? m1 : a[k]));
      m1 = g ? a[k] : m1;
      i1 = g ? ix[k] : i1;
    }
    const bool banded = (m1 < 8.0f) && (m1 > -80.0f);
    lo = banded ? (m1 - 0.01f) : -INFINITY;
    slow = !(nan_acc == 0.0f) || !(r2 < lo) || i1 >= C;
    best_s = m1;  // sigmoid applied below
    best_c = i1;
  }
  if (want && !slow) best_s = dm_sigmoidf(best_s);
  // cooperative exact path, one record at a time
  uint32_t todo = __ballot_sync(0xffffffffu, want && slow);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1u;
    const float lo_s = __shfl_sync(0xffffffffu, lo, src);
    const float* __restrict__ rc = slab + src * RF + 5;
    // serial semantics being reproduced: walk c upwards over the in-band logits (x >= lo or NaN); the first one
    // initialises (s, c) even when its sigmoid is NaN; later ones replace it only when s > best
    float ls = 0.0f;
    int lc = 0x7fffffff, lfirst = 0x7fffffff;
    bool lfirst_nan = false;
    for (int c = lane; c < C; c += 32) {
      const float x = rc[c];
      if (x >= lo_s || x != x) {
        const float sg = dm_sigmoidf(x);
        if (lfirst == 0x7fffffff) { lfirst = c; lfirst_nan = sg != sg; }
        if (sg == sg && (lc == 0x7fffffff || sg > ls)) { ls = sg; lc = c; }
      }
    }
    int first = lfirst;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    const bool first_is_nan = __any_sync(0xffffffffu, lfirst == first && first != 0x7fffffff && lfirst_nan);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ls, o);
      const int oc = __shfl_xor_sync(0xffffffffu, lc, o);
      const bool take = (oc != 0x7fffffff) && (lc == 0x7fffffff || os > ls || (os == ls && oc < lc));
      ls = take ? os : ls;
      lc = take ? oc : lc;
    }
    if (lane == src) {
      if (first == 0x7fffffff) { best_s = 0.0f; best_c = 0; }               // no in-band logit at all (cannot happen for C >= 1)
      else if (first_is_nan || lc == 0x7fffffff) { best_s = __int_as_float(0x7fc00000); best_c = first; }
      else { best_s = ls; best_c = lc; }
    }
  }
}

struct Decoded { float x1, y1, x2, y2; bool valid; };

__device__ __forceinline__ Decoded decode_box(float tx, float ty, float tw, float th, int gx, int gy, int W, int H,
                                              float aw, float ah) {
  // tyu:153-161: xy = (sigmoid(t)+grid)/grid_wh; wh = exp(t)*anchor (anchor already /image_wh); inf -> 0
  Decoded d;
  float x = DM_DIV(DM_ADD(dm_sigmoidf(tx), (float)gx), (float)W);
  float y = DM_DIV(DM_ADD(dm_sigmoidf(ty), (float)gy), (float)H);
  float w = DM_MUL(dm_expf(tw), aw);
  float h = DM_MUL(dm_expf(th), ah);
  if (dm_isinf(w)) w = 0.0f;
  if (dm_isinf(h)) h = 0.0f;
  float hw = DM_DIV(w, 2.0f), hh = DM_DIV(h, 2.0f);
  d.x1 = DM_SUB(x, hw); d.y1 = DM_SUB(y, hh); d.x2 = DM_ADD(x, hw); d.y2 = DM_ADD(y, hh);
  d.valid = (d.x2 > d.x1) && (d.y2 > d.y1);
  return d;
}

template <bool EARLY>   // EARLY: a warp refills its slab before its append (atomic + stores) instead of after it
__global__ void __launch_bounds__(512, 1) yolo_decode_filter_kernel(YoloDecodeParams p) {
  extern __shared__ __align__(128) unsigned char yd_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int RF = p.RF;
  const uint32_t slab_bytes = 128u * (uint32_t)RF;
  // layout: [warps][stages] slabs, then barriers
  float* slabs = reinterpret_cast<float*>(yd_smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(yd_smem + (size_t)warps_per_cta * YD_STAGES * slab_bytes);
  float* const slab0 = slabs + (size_t)(warp * YD_STAGES) * (slab_bytes / 4);
  uint64_t* const bar0 = bars + warp * YD_STAGES;
#define my_slab(s) (slab0 + (size_t)(s) * (slab_bytes / 4))
#define my_bar(s) (bar0 + (s))
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < YD_STAGES; ++s) mbar_init(my_bar(s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const long long n_tiles = p.lv.tile_base[YD_MAX_LEVELS];
  const long long gwarp = (long long)blockIdx.x * warps_per_cta + warp;
  const long long gstride = (long long)gridDim.x * warps_per_cta;

  // issue the load of `tile` into stage s
  auto issue = [&](long long tile, int s) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < YD_MAX_LEVELS; ++k) if (tile >= p.lv.tile_base[k]) l = k;
    const long long rec0 = (tile - p.lv.tile_base[l]) * 32;
    const long long remain = p.lv.total_rec[l] - rec0;
    const int nrec = remain < 32 ? (int)remain : 32;
    const float* src = p.lv.head[l] + rec0 * RF;
    const uint32_t bytes = (uint32_t)nrec * (uint32_t)RF * 4u;
    if ((bytes & 15u) == 0u) {
      if (lane == 0) {
        mbar_expect_tx(my_bar(s), bytes);
        bulk_g2s(my_slab(s), src, bytes, my_bar(s));
      }
    } else {  // ragged tail of a level: plain copy, then a transaction-less arrive flips the same phase
      for (int i = lane; i < nrec * RF; i += 32) my_slab(s)[i] = __ldg(src + i);
      __syncwarp();
      if (lane == 0) mbar_arrive(my_bar(s));
    }
  };

  long long t_issue = gwarp;
#pragma unroll
  for (int s = 0; s < YD_STAGES; ++s) {
    if (t_issue < n_tiles) issue(t_issue, s);
    t_issue += gstride;
  }
  uint32_t phase = 0;
  int stage = 0;
  for (long long tile = gwarp; tile < n_tiles; tile += gstride) {
    mbar_wait(my_bar(stage), phase);
    int l = 0;
#pragma unroll
    for (int k = 1; k < YD_MAX_LEVELS; ++k) if (tile >= p.lv.tile_base[k]) l = k;
    const long long rec = (tile - p.lv.tile_base[l]) * 32 + lane;
    const bool in_range = rec < p.lv.total_rec[l];
    const float* r = my_slab(stage) + lane * RF;
    bool pass = false;
    int img = 0;
    float score = 0.f, conf = 0.f;
    int cls = 0;
    uint32_t aidx = 0;
    bool want = false;
    if (in_range) {
      // sigmoid(conf) > conf_thr (tyu:191), decided in logit space outside a narrow band around logit(conf_thr)
      conf = r[4];
      if (conf > p.conf_hi) want = true;
      else if (conf >= p.conf_lo) want = dm_sigmoidf(conf) > p.conf_thr;  // NaN fails both comparisons: not wanted
    }
    class_max_sigmoid_warp(my_slab(stage), RF, p.C, want, lane, score, cls);
    if (want && score > p.score_thr) {
      const int rpi = p.lv.rec_per_img[l];
      // 32-bit division when the level has fewer than 2^31 records (always, short of ~100k-image batches)
      img = (p.lv.total_rec[l] < 0x7fffffffLL) ? (int)((uint32_t)rec / (uint32_t)rpi) : (int)(rec / rpi);
      const int rin = (int)(rec - (long long)img * rpi);
      const int cell = p.magic_a ? (int)__umulhi((uint32_t)rin, p.magic_a) : rin, a = rin - cell * p.A;  // rin / A
      // the box itself is decoded by the NMS pass, and only for the candidates it looks at; here only its validity
      // (x2 > x1 and y2 > y1, tyu:163) is needed: certain inside the safe logit range, exact decode otherwise
      const float tx = r[0], ty = r[1], tw = r[2], th = r[3];
      bool valid = (tw >= p.lv.tmin_w[l][a]) && (tw <= 80.0f) && (th >= p.lv.tmin_h[l][a]) && (th <= 80.0f) && (tx == tx) && (ty == ty);
      if (!valid) {
        const int W = p.lv.w[l], H = p.lv.h[l];
        const int gy = cell / W, gx = cell - gy * W;
        valid = decode_box(tx, ty, tw, th, gx, gy, W, H, p.lv.anc_w[l][a], p.lv.anc_h[l][a]).valid;
      }
      if (valid) {
        pass = true;
        aidx = (uint32_t)(p.lv.anchor_base[l] + rin);
      }
    }
    // Every lane is done reading the slab: refill it NOW, so that the copy of the next tile is in flight during the append
    // below (an atomic round trip to L2 plus the stores — 9 % of the warp's cycle when the copy was issued after it).
    if (EARLY) {
      __syncwarp();
      if (t_issue < n_tiles) issue(t_issue, stage);
      t_issue += gstride;
    }
    // warp-aggregated append, one atomic per (warp, image)
    uint32_t todo = __ballot_sync(0xffffffffu, pass);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int limg = __shfl_sync(0xffffffffu, img, leader);
      const uint32_t grp = __ballot_sync(0xffffffffu, pass && img == limg);
      int base = 0;
      if (lane == leader) base = atomicAdd(&p.counts[limg], __popc(grp));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (pass && img == limg) {
        const size_t slot = (size_t)limg * p.n_img + base + __popc(grp & ((1u << lane) - 1u));
        p.cand_score[slot] = score;
        p.cand_cls[slot] = cls;
        p.cand_conf[slot] = conf;  // the logit; the sigmoid is applied to the emitted rows only
        p.cand_aidx[slot] = aidx;
        if (p.bitmap) atomicOr(&p.bitmap[(size_t)limg * p.bitmap_words + (aidx >> 5)], 1u << (aidx & 31u));
      }
      todo &= ~grp;
    }
    if (!EARLY) {
      __syncwarp();  // every lane is done reading the slab before it is refilled
      if (t_issue < n_tiles) issue(t_issue, stage);
      t_issue += gstride;
    }
    if (++stage == YD_STAGES) { stage = 0; phase ^= 1u; }
  }
}

struct YoloFinalizeParams {
  YoloLevels lv;
  int B, A, C, RF, n_img;
  NmsConfig cfg;
  float4* cand_box;  // written by the NMS pass (decode on demand)
  const float* cand_score; const int32_t* cand_cls; const float* cand_conf;
  const uint32_t* cand_aidx; const int32_t* counts; const uint32_t* bitmap; int bitmap_words;
  int32_t* nms_pos;  // [B, max_out] scratch
  // outputs, all [B, max_out, ...]
  float* out_boxes; int32_t* out_cls; float* out_score; float* out_classes; float* out_conf;
  int32_t* out_sel_idx; int32_t* out_sel_anchor; int32_t* out_count;
};

// Decode-on-demand for NMS: candidate `pos` of the image -> its record in the head tensor -> decode_box; the result is
// written back to the candidate store so that the emitted rows can be copied out afterwards.
struct YoloLazyBox {
  static constexpr bool kKeyCache = true;   // a 416x416 image has ~5 k candidates: its selection passes run from registers
  static constexpr int kMode = B200_NMS_BY_CLASS;
  static constexpr bool kSampleRange = false;
  const YoloLevels* lv;
  int img, A, RF;
  float4* cand_box;  // this image's slice
  __device__ __forceinline__ float4 operator()(const NmsSegment& seg, uint32_t pos) const {
    const int a_flat = (int)seg.order_id[pos];
    int l = 0;
#pragma unroll
    for (int j = 1; j < YD_MAX_LEVELS; ++j) if (a_flat >= lv->anchor_base[j]) l = j;
    const int rin = a_flat - lv->anchor_base[l];
    const float* r = lv->head[l] + ((long long)img * lv->rec_per_img[l] + rin) * RF;
    const int cell = rin / A, a = rin - cell * A;
    const int W = lv->w[l], H = lv->h[l];
    const int gy = cell / W, gx = cell - gy * W;
    const Decoded d = decode_box(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3), gx, gy, W, H, lv->anc_w[l][a], lv->anc_h[l][a]);
    const float4 b = make_float4(d.x1, d.y1, d.x2, d.y2);
    cand_box[pos] = b;
    return b;
  }
};

template <int METRIC, int THREADS>
__global__ void __launch_bounds__(THREADS, NMS_THREADS / THREADS) yolo_nms_finalize_kernel(YoloFinalizeParams p) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  NMS_T(13);
  const int img = blockIdx.x;
  const size_t cbase = (size_t)img * p.n_img;
  NmsSegment seg;
  seg.boxes = reinterpret_cast<const float*>(p.cand_box + cbase);
  seg.scores = p.cand_score + cbase;
  seg.classes = p.cand_cls + cbase;
  seg.order_id = p.cand_aidx + cbase;
  seg.n = p.counts[img];
  int32_t* pos = p.nms_pos + (size_t)img * p.cfg.max_out;
  YoloLazyBox lazy;
  lazy.lv = &p.lv; lazy.img = img; lazy.A = p.A; lazy.RF = p.RF; lazy.cand_box = p.cand_box + cbase;
  const int kept = nms_run_segment<METRIC, YoloLazyBox, THREADS>(seg, p.cfg, pos, nms_smem, nullptr, lazy);
  __syncthreads();
  NMS_T(11);
  if (threadIdx.x == 0) p.out_count[img] = kept;
  const size_t obase = (size_t)img * p.cfg.max_out;
  // rank of an anchor index among the image's candidates = its position in the reference's compacted list
  uint32_t* wprefix = reinterpret_cast<uint32_t*>(nms_smem);  // reuse (NMS is finished)
  if (p.out_sel_idx && p.bitmap) {
    const uint32_t* bm = p.bitmap + (size_t)img * p.bitmap_words;
    // exclusive prefix popcount over words, single pass by warp 0 (<= a few thousand words)
    if (threadIdx.x < 32) {
      uint32_t run = 0;
      for (int w0 = 0; w0 < p.bitmap_words; w0 += 32) {
        int w = w0 + (int)threadIdx.x;
        int c = (w < p.bitmap_words) ? __popc(bm[w]) : 0;
        int inc = warp_scan_incl(c);
        if (w < p.bitmap_words) wprefix[w] = run + (uint32_t)(inc - c);
        run += (uint32_t)__shfl_sync(0xffffffffu, inc, 31);
      }
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < kept; k += blockDim.x) {
    const int q = pos[k];
    const float4 b = p.cand_box[cbase + q];
    reinterpret_cast<float4*>(p.out_boxes)[obase + k] = b;
    p.out_cls[obase + k] = p.cand_cls[cbase + q];
    p.out_score[obase + k] = p.cand_score[cbase + q];
    p.out_conf[obase + k] = dm_sigmoidf(p.cand_conf[cbase + q]);
    const uint32_t a = p.cand_aidx[cbase + q];
    pos[k] = (int32_t)a;   // from here on the scratch row holds flat anchor indices: yolo_classes_kernel needs no candidate lookup
    if (p.out_sel_anchor) p.out_sel_anchor[obase + k] = (int32_t)a;
    if (p.out_sel_idx && p.bitmap) {
      const uint32_t wv = p.bitmap[(size_t)img * p.bitmap_words + (a >> 5)];
      p.out_sel_idx[obase + k] = (int32_t)(wprefix[a >> 5] + __popc(wv & ((1u << (a & 31u)) - 1u)));
    }
  }
  NMS_T(12);
}

#ifdef NMS_TRACE
extern "C" int b200_debug_nms_trace(long long* out_host64) {
  return cudaMemcpyFromSymbol(out_host64, g_nms_trace, sizeof(long long) * (64 + 4 * 32)) == cudaSuccess ? 0 : -1;
}
#endif

// sigmoid(classes) rows of the selected boxes, read back from the head tensors (tyu:140,265).  A separate launch
// so the B*max_out*C sigmoids spread over the whole GPU instead of serialising inside the per-image NMS CTAs.
// LPR lanes per row: the kernel is bound by the issue of the deterministic sigmoid (~60 instructions), so the lanes that
// idle in a row's last round are the cost — 80 classes on 32 lanes waste one round in six, on 16 lanes (two rows per
// warp) none.
template <int LPR>
__global__ void __launch_bounds__(256) yolo_classes_kernel(YoloFinalizeParams p) {
  constexpr int RPW = 32 / LPR, NJ = 128 / LPR;
  const int sub = threadIdx.x & (LPR - 1);
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + ((threadIdx.x & 31) / LPR);  // row = img * max_out + k
  if (row >= p.B * p.cfg.max_out) return;
  const int img = row / p.cfg.max_out, k = row - img * p.cfg.max_out;
  // flat anchor index, left by the NMS kernel; fetched together with the count (rows past the count hold stale values)
  const uint32_t a = (uint32_t)__ldcg(p.nms_pos + (size_t)img * p.cfg.max_out + k);
  if (k >= __ldcg(p.out_count + img)) return;
  int l = 0;
#pragma unroll
  for (int j = 1; j < YD_MAX_LEVELS; ++j) if ((int)a >= p.lv.anchor_base[j]) l = j;
  const long long rec = (long long)img * p.lv.rec_per_img[l] + ((int)a - p.lv.anchor_base[l]);
  const float* src = p.lv.head[l] + rec * p.RF + 5;
  float* dst = p.out_classes + (size_t)row * p.C;
  // the (cold) logits of up to 128 classes are fetched before the first sigmoid starts
  float x[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) x[j] = (sub + LPR * j < p.C) ? __ldg(src + sub + LPR * j) : 0.0f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) if (sub + LPR * j < p.C) dst[sub + LPR * j] = dm_sigmoidf(x[j]);
  for (int c = sub + 128; c < p.C; c += LPR) dst[c] = dm_sigmoidf(__ldg(src + c));
}

// ---- dense decode for the stand-alone GetBoxes shim ---------------------------------------------
// One warp per record: boxes/conf/sigmoid(classes)/valid for every anchor of one level, no filtering.
__global__ void yolo_decode_dense_kernel(const float* __restrict__ head, long long total_rec, int H, int W, int A,
                                         int C,
                                         const float* __restrict__ anc_wh, float* __restrict__ boxes,
                                         float* __restrict__ conf, float* __restrict__ classes,
                                         unsigned char* __restrict__ valid) {
  const int RF = 5 + C;
  const int lane = threadIdx.x & 31;
  long long rec = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
  const int rpi = H * W * A;
  for (; rec < total_rec; rec += stride) {
    const float* r = head + rec * RF;
    for (int c = lane; c < C; c += 32) classes[rec * C + c] = dm_sigmoidf(__ldg(r + 5 + c));
    if (lane == 0) {
      const int rin = (int)(rec % rpi);
      const int cell = rin / A, a = rin - cell * A;
      const int gy = cell / W, gx = cell - gy * W;
      Decoded d = decode_box(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3), gx, gy, W, H, anc_wh[2 * a], anc_wh[2 * a + 1]);
      boxes[rec * 4 + 0] = d.x1; boxes[rec * 4 + 1] = d.y1; boxes[rec * 4 + 2] = d.x2; boxes[rec * 4 + 3] = d.y2;
      conf[rec] = dm_sigmoidf(__ldg(r + 4));
      valid[rec] = d.valid ? 1 : 0;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------
static int fill_levels(YoloLevels& lv, const float* const heads[3], const int32_t hw[6], int B, int A, int C,
                       const float* anchors_wh_host, const float* image_wh_host) {
  long long tb = 0;
  int ab = 0;
  for (int l = 0; l < YD_MAX_LEVELS; ++l) {
    lv.head[l] = heads[l];
    lv.h[l] = hw[2 * l];
    lv.w[l] = hw[2 * l + 1];
    lv.rec_per_img[l] = lv.h[l] * lv.w[l] * A;
    lv.anchor_base[l] = ab;
    ab += lv.rec_per_img[l];
    lv.total_rec[l] = (long long)B * lv.rec_per_img[l];
    lv.tile_base[l] = tb;
    tb += (lv.total_rec[l] + 31) / 32;
    for (int a = 0; a < A; ++a) {
      lv.anc_w[l][a] = anchors_wh_host[(l * A + a) * 2 + 0] / image_wh_host[0];
      lv.anc_h[l][a] = anchors_wh_host[(l * A + a) * 2 + 1] / image_wh_host[1];
      // w = exp(tw) * anchor >= 4e-7 (twice the 2e-7 that guarantees x + w/2 > x - w/2 for x in [0,1]); anchors that
      // are not positive and finite never qualify
      const double aw = lv.anc_w[l][a], ah = lv.anc_h[l][a];
      lv.tmin_w[l][a] = (aw > 0.0 && aw < 1e30) ? (float)(log(4e-7 / aw) + 1e-3) : INFINITY;
      lv.tmin_h[l][a] = (ah > 0.0 && ah < 1e30) ? (float)(log(4e-7 / ah) + 1e-3) : INFINITY;
    }
  }
  lv.tile_base[YD_MAX_LEVELS] = tb;
  return ab;
}

struct YoloWs {
  size_t counts, bitmap, box, score, cls, conf, aidx, pos, total;
  int bitmap_words;
};
static YoloWs yolo_ws_layout(int B, int n_img, int max_out) {
  YoloWs w;
  size_t o = 0;
  w.bitmap_words = (n_img + 31) / 32;
  w.counts = o; o = b200_align_up(o + sizeof(int32_t) * B, 256);
  w.bitmap = o; o = b200_align_up(o + sizeof(uint32_t) * (size_t)B * w.bitmap_words, 256);
  w.box = o; o = b200_align_up(o + sizeof(float4) * (size_t)B * n_img, 256);
  w.score = o; o = b200_align_up(o + sizeof(float) * (size_t)B * n_img, 256);
  w.cls = o; o = b200_align_up(o + sizeof(int32_t) * (size_t)B * n_img, 256);
  w.conf = o; o = b200_align_up(o + sizeof(float) * (size_t)B * n_img, 256);
  w.aidx = o; o = b200_align_up(o + sizeof(uint32_t) * (size_t)B * n_img, 256);
  w.pos = o; o = b200_align_up(o + sizeof(int32_t) * (size_t)B * max_out, 256);
  w.total = o;
  return w;
}

extern "C" size_t b200_yolo_decode_nms_workspace_bytes(const int32_t hw[6], int B, int A, int max_out) {
  int n_img = 0;
  for (int l = 0; l < 3; ++l) n_img += hw[2 * l] * hw[2 * l + 1] * A;
  return yolo_ws_layout(B, n_img, max_out).total;
}

extern "C" int b200_yolo_decode_nms(const float* const heads[3], const int32_t hw[6], int B, int A, int C,
                                    const float* anchors_wh_host, const float* image_wh_host, float conf_thr,
                                    float score_thr, float iou_thr, int metric, int max_out, float* out_boxes,
                                    int32_t* out_class_id, float* out_score, float* out_classes, float* out_conf,
                                    int32_t* out_sel_idx, int32_t* out_sel_anchor, int32_t* out_count,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_REQUIRE(heads && hw && anchors_wh_host && image_wh_host, B200_ERR_BAD_ARG, "b200_yolo_decode_nms: null argument");
  B200_REQUIRE(B >= 0 && A >= 1 && A <= 8 && C >= 1 && C <= 1019, B200_ERR_BAD_ARG,
               "b200_yolo_decode_nms: unsupported shape B=%d A=%d C=%d", B, A, C);
  B200_REQUIRE(metric >= B200_METRIC_YOLO_IOU && metric <= B200_METRIC_YOLO_CIOU, B200_ERR_BAD_ARG,
               "b200_yolo_decode_nms: iou_type must be iou/diou/ciou (metric %d)", metric);
  B200_REQUIRE(max_out >= 1 && max_out <= NMS_MAX_OUT_LIMIT, B200_ERR_UNSUPPORTED, "b200_yolo_decode_nms: max_out %d outside [1,%d]", max_out, NMS_MAX_OUT_LIMIT);
  if (B == 0) return B200_OK;
  B200_REQUIRE(out_boxes && out_class_id && out_score && out_conf && out_count, B200_ERR_BAD_ARG, "b200_yolo_decode_nms: null output");
  for (int l = 0; l < 3; ++l) {
    B200_REQUIRE(heads[l] != nullptr && hw[2 * l] > 0 && hw[2 * l + 1] > 0, B200_ERR_BAD_ARG, "b200_yolo_decode_nms: bad level %d", l);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(heads[l]) & 15) == 0, B200_ERR_BAD_ARG, "b200_yolo_decode_nms: head %d not 16-byte aligned", l);
  }
  B200_REQUIRE((reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0, B200_ERR_BAD_ARG, "b200_yolo_decode_nms: out_boxes not 16-byte aligned");
  YoloDecodeParams dp;
  const int n_img = fill_levels(dp.lv, heads, hw, B, A, C, anchors_wh_host, image_wh_host);
  YoloWs ws = yolo_ws_layout(B, n_img, max_out);
  B200_REQUIRE(workspace && workspace_bytes >= ws.total, B200_ERR_WORKSPACE,
               "b200_yolo_decode_nms: workspace %zu < required %zu", workspace_bytes, ws.total);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, B200_ERR_BAD_ARG, "b200_yolo_decode_nms: workspace not 256-byte aligned");
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  dp.B = B; dp.A = A; dp.C = C; dp.RF = 5 + C; dp.n_img = n_img;
  dp.conf_thr = conf_thr; dp.score_thr = score_thr;
  dp.magic_a = A == 1 ? 0u : (uint32_t)((1ull << 32) / (unsigned long long)A + 1ull);
  // sigmoid(x) > conf_thr is certain for x > logit(thr) + d and certainly false for x < logit(thr) - d, with d ten times
  // the 2.4-ulp error of the deterministic sigmoid divided by the slope thr(1-thr); thresholds near 0 or 1 (or outside
  // (0,1)) always take the exact comparison
  dp.conf_lo = -INFINITY; dp.conf_hi = INFINITY;
  if (conf_thr >= 1e-3f && conf_thr <= 1.0f - 1e-3f) {
    const double t = log((double)conf_thr / (1.0 - (double)conf_thr));
    const double d = 3e-6 / ((double)conf_thr * (1.0 - (double)conf_thr)) + 1e-6 * fabs(t);
    dp.conf_lo = (float)(t - d); dp.conf_hi = (float)(t + d);
  }
  dp.cand_box = reinterpret_cast<float4*>(wsb + ws.box);
  dp.cand_score = reinterpret_cast<float*>(wsb + ws.score);
  dp.cand_cls = reinterpret_cast<int32_t*>(wsb + ws.cls);
  dp.cand_conf = reinterpret_cast<float*>(wsb + ws.conf);
  dp.cand_aidx = reinterpret_cast<uint32_t*>(wsb + ws.aidx);
  dp.counts = reinterpret_cast<int32_t*>(wsb + ws.counts);
  dp.bitmap = out_sel_idx ? reinterpret_cast<uint32_t*>(wsb + ws.bitmap) : nullptr;
  dp.bitmap_words = ws.bitmap_words;
  // counts and bitmap are adjacent at the front of the workspace: one memset
  B200_CUDA(cudaMemsetAsync(wsb, 0, out_sel_idx ? ws.box : ws.bitmap, stream));

  const uint32_t slab = 128u * (uint32_t)dp.RF;
  int warps = 16;  // one tile in flight per warp; the warps of an SM overlap each other's loads
  while (warps > 1 && (size_t)warps * YD_STAGES * slab + 256 > 200 * 1024) warps >>= 1;
  // a single image has fewer tiles than the GPU has warp slots: spread them (one warp per scheduler finishes its tile sooner)
  while (warps > 2 && dp.lv.tile_base[YD_MAX_LEVELS] <= (long long)(warps / 2) * b200_sm_count()) warps >>= 1;
  B200_REQUIRE((size_t)warps * YD_STAGES * slab + 256 <= 220 * 1024, B200_ERR_UNSUPPORTED, "b200_yolo_decode_nms: record too large (C=%d)", C);
  const size_t smem1 = (size_t)warps * YD_STAGES * slab + sizeof(uint64_t) * warps * YD_STAGES + 16;

  const long long n_tiles = dp.lv.tile_base[YD_MAX_LEVELS];
  long long want = (n_tiles + warps - 1) / warps;
  int grid = (int)(want < (long long)b200_sm_count() ? want : (long long)b200_sm_count());
  if (grid < 1) grid = 1;
  // Refilling a warp's slab before its append keeps one more copy in flight: -3 % for the call on its own at any size, and
  // for the fused evaluation step at 512 images; at 64 images per GPU the same step is 3 % SLOWER with it (the filter then
  // takes DRAM bandwidth from the loss chain that shares the GPU, and that chain is the longer one there).  Long streams
  // (>= 32 tiles per warp) refill early, short ones late; B200_YD_EARLY_ISSUE=0/1 overrides (measurements only).
  { const char* e = getenv("B200_YD_EARLY_ISSUE"); dp.early_issue = e ? atoi(e) : (n_tiles >= 32ll * grid * warps ? 1 : 0); }
  if (dp.early_issue) {
    B200_CUDA(cudaFuncSetAttribute(yolo_decode_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    yolo_decode_filter_kernel<true><<<grid, warps * 32, smem1, stream>>>(dp);
  } else {
    B200_CUDA(cudaFuncSetAttribute(yolo_decode_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    yolo_decode_filter_kernel<false><<<grid, warps * 32, smem1, stream>>>(dp);
  }
  B200_LAUNCH_CHECK();

  YoloFinalizeParams fp;
  fp.lv = dp.lv; fp.B = B; fp.A = A; fp.C = C; fp.RF = dp.RF; fp.n_img = n_img;
  fp.cfg.metric = metric; fp.cfg.mode = B200_NMS_BY_CLASS; fp.cfg.iou_thr = iou_thr; fp.cfg.score_thr = 0.f;
  fp.cfg.use_score_thr = 0; fp.cfg.max_out = max_out;
  fp.cand_box = dp.cand_box; fp.cand_score = dp.cand_score; fp.cand_cls = dp.cand_cls; fp.cand_conf = dp.cand_conf;
  fp.cand_aidx = dp.cand_aidx; fp.counts = dp.counts; fp.bitmap = dp.bitmap; fp.bitmap_words = ws.bitmap_words;
  fp.nms_pos = reinterpret_cast<int32_t*>(wsb + ws.pos);
  fp.out_boxes = out_boxes; fp.out_cls = out_class_id; fp.out_score = out_score; fp.out_classes = out_classes;
  fp.out_conf = out_conf; fp.out_sel_idx = out_sel_idx; fp.out_sel_anchor = out_sel_anchor; fp.out_count = out_count;
  // one 1024-thread CTA per SM gives the lowest latency; batches that need more than one wave use 512-thread CTAs,
  // two per SM (the shared-memory footprint allows it up to max_out ~ 500)
  size_t smem2 = nms_smem_bytes(max_out);
  const size_t need_prefix = (size_t)ws.bitmap_words * 4;
  if (smem2 < need_prefix) smem2 = need_prefix;
  const bool half_ctas = (B > b200_sm_count()) && (2 * (smem2 + 1024) <= 227 * 1024);
#define YD_LAUNCH(M)                                                                                                   \
  if (half_ctas) {                                                                                                          \
    B200_CUDA(cudaFuncSetAttribute(yolo_nms_finalize_kernel<M, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
    yolo_nms_finalize_kernel<M, 512><<<B, 512, smem2, stream>>>(fp);                                                         \
  } else {                                                                                                                  \
    B200_CUDA(cudaFuncSetAttribute(yolo_nms_finalize_kernel<M, NMS_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
    yolo_nms_finalize_kernel<M, NMS_THREADS><<<B, NMS_THREADS, smem2, stream>>>(fp);                                         \
  }
  NMS_DISPATCH_METRIC(metric, YD_LAUNCH)
#undef YD_LAUNCH
  B200_LAUNCH_CHECK();
  if (out_classes) {
    const long long rows = (long long)B * max_out;
    // half a warp per row when that leaves fewer idle lanes in the last round (80 classes: 5 x 16 against 3 x 32)
    // (a few hundred rows are latency-bound: a full warp per row has the shorter chain)
    if (rows >= 16384 && ((C + 15) / 16) * 16 < ((C + 31) / 32) * 32) yolo_classes_kernel<16><<<(int)((rows + 15) / 16), 256, 0, stream>>>(fp);
    else yolo_classes_kernel<32><<<(int)((rows + 7) / 8), 256, 0, stream>>>(fp);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

extern "C" int b200_yolo_decode_dense(const float* head, int B, int H, int W, int A, int C,
                                      const float* anchors_wh_norm_dev, float* boxes, float* conf, float* classes,
                                      unsigned char* valid, void* stream) {
  B200_REQUIRE(B >= 0 && H > 0 && W > 0 && A >= 1 && C >= 1, B200_ERR_BAD_ARG, "b200_yolo_decode_dense: bad shape");
  if (B == 0) return B200_OK;
  B200_REQUIRE(head && anchors_wh_norm_dev && boxes && conf && classes && valid, B200_ERR_BAD_ARG, "b200_yolo_decode_dense: null pointer");
  const long long total = (long long)B * H * W * A;
  long long blocks = (total + 7) / 8;
  long long cap = (long long)b200_sm_count() * 64;
  if (blocks > cap) blocks = cap;
  yolo_decode_dense_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(head, total, H, W, A, C, anchors_wh_norm_dev, boxes, conf,
                                                                         classes, valid);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
