// effdet.cu — EfficientDet head utilities: anchors, box decode, per-image post-processing, target assignment.
//
// Replaces (paths relative to AIServer/ai_api/ai_models/):
//   efficientnet/utils/anchors.py:46-84    Anchors._generate_boxes        -> effdet_anchors_kernel
//   efficientnet/utils/anchors.py:141-158, 245-274  convert_outputs_boxes / _boxes_decoder -> effdet_decode_kernel
//   efficientnet/utils/anchors.py:161-202  convert_outputs_one            -> effdet_filter_kernel + effdet_nms_finalize_kernel
//   efficientnet/utils/anchors.py:91-138, 219-243   generate_targets / _boxes_encoder -> effdet_assign_kernel
//
// Anchors are never read from HBM: every kernel rebuilds an anchor box from a tiny per-level table
// (row centres yc[H], column centres xc[W], half extents hy[A], hx[A]; built on the host with the reference's own
// double-precision Python arithmetic) exactly as the reference does — box = [yc-hy, xc-hx, yc+hy, xc+hx] in fp32,
// then centre/size re-derived from those corners (anc:204-217) — so results are bit-identical to reading
// Anchors.boxes while saving 0.8 MB (D0) / 7 MB (D7) of reads per image.
#include "nms.cuh"
#include "effdet_focal.cuh"
#include "bulkcopy.cuh"

#define EF_MAX_LEVELS 8
#define EF_STAGES 2

struct EfLevels {
  int num_levels, A;
  int h[EF_MAX_LEVELS], w[EF_MAX_LEVELS];
  int anc_per_img[EF_MAX_LEVELS];   // h*w*A
  int anchor_base[EF_MAX_LEVELS];   // flat anchor index of the level's first anchor within an image
  const float* table;               // device
  int tab_off[EF_MAX_LEVELS];       // float offset of the level's [yc(h) | xc(w) | hy(A) | hx(A)] block
};

struct AnchorBox { float y1, x1, y2, x2; };

__device__ __forceinline__ AnchorBox ef_anchor(const EfLevels& lv, int l, int y, int x, int a) {
  const float* t = lv.table + lv.tab_off[l];
  const float yc = __ldg(t + y), xc = __ldg(t + lv.h[l] + x);
  const float hy = __ldg(t + lv.h[l] + lv.w[l] + a), hx = __ldg(t + lv.h[l] + lv.w[l] + lv.A + a);
  AnchorBox b;
  b.y1 = DM_SUB(yc, hy); b.x1 = DM_SUB(xc, hx); b.y2 = DM_ADD(yc, hy); b.x2 = DM_ADD(xc, hx);  // anc:77-78
  return b;
}

__device__ __forceinline__ void ef_split(const EfLevels& lv, int l, int rin, int& y, int& x, int& a) {
  const int cell = rin / lv.A;
  a = rin - cell * lv.A;
  y = cell / lv.w[l];
  x = cell - y * lv.w[l];
}

// ---- anchors -------------------------------------------------------------------------------------
__global__ void effdet_anchors_kernel(EfLevels lv, int l, float4* __restrict__ out) {
  const int n = lv.anc_per_img[l];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int y, x, a;
    ef_split(lv, l, i, y, x, a);
    const AnchorBox b = ef_anchor(lv, l, y, x, a);
    out[i] = make_float4(b.y1, b.x1, b.y2, b.x2);
  }
}

// ---- decode (anc:245-274) ---------------------------------------------------------------------------
struct EfDecodeParams {
  EfLevels lv;
  int B;
  const float4* rel[EF_MAX_LEVELS];
  float4* out[EF_MAX_LEVELS];
  long long elem_base[EF_MAX_LEVELS + 1];  // cumulative B*anc_per_img
};

__global__ void __launch_bounds__(256) effdet_decode_kernel(EfDecodeParams p) {
  const long long total = p.elem_base[p.lv.num_levels];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < EF_MAX_LEVELS; ++k) if (k < p.lv.num_levels && i >= p.elem_base[k]) l = k;
    const long long e = i - p.elem_base[l];
    const unsigned int api = (unsigned int)p.lv.anc_per_img[l];
    const unsigned int img = (unsigned int)(e / api);  // e < 2^31*api in practice; 64/32 division only here
    const int rin = (int)(e - (long long)img * api);
    int y, x, a;
    ef_split(p.lv, l, rin, y, x, a);
    const AnchorBox an = ef_anchor(p.lv, l, y, x, a);
    const float yca = DM_DIV(DM_ADD(an.y2, an.y1), 2.0f), xca = DM_DIV(DM_ADD(an.x2, an.x1), 2.0f);
    const float ha = DM_SUB(an.y2, an.y1), wa = DM_SUB(an.x2, an.x1);
    const float4 r = __ldcs(p.rel[l] + e);  // ty, tx, th, tw
    const float w = DM_MUL(dm_expf(r.w), wa), h = DM_MUL(dm_expf(r.z), ha);
    const float yc = DM_ADD(DM_MUL(r.x, ha), yca), xc = DM_ADD(DM_MUL(r.y, wa), xca);
    const float hh = DM_DIV(h, 2.0f), hw = DM_DIV(w, 2.0f);
    __stcs(p.out[l] + e, make_float4(DM_SUB(yc, hh), DM_SUB(xc, hw), DM_ADD(yc, hh), DM_ADD(xc, hw)));
  }
}

// ---- class argmax / background filter (anc:168-189), streaming like yolo_decode_filter_kernel ----------
struct EfFilterParams {
  EfLevels lv;
  int B0, NB, C;                 // images [B0, B0+NB) of a batch of size B
  int B;
  int n_img;                     // anchors per image
  const float* cls[EF_MAX_LEVELS];    // (B,H,W,A,C) logits
  const float4* boxes[EF_MAX_LEVELS]; // (B,H,W,A,4) decoded, or null: decode here from rel (lv.table must be set)
  const float4* rel[EF_MAX_LEVELS];   // (B,H,W,A,4) head offsets ty,tx,th,tw (fused decode mode)
  float4* dec[EF_MAX_LEVELS];         // fused decode mode: the dense decoded tensor is written here (may be null)
  long long tile_base[EF_MAX_LEVELS + 1];  // tiles of 32 anchors over the NB images of each level
  float4* cand_box; float* cand_score; int32_t* cand_cls; uint32_t* cand_aidx;  // [NB, n_img]
  int32_t* counts;               // [NB]
  uint32_t* bitmap; int bitmap_words;
};

__device__ __forceinline__ uint32_t ef_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) effdet_filter_kernel(EfFilterParams p) {
  extern __shared__ __align__(128) unsigned char ef_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int C = p.C;
  const uint32_t slab_bytes = 128u * (uint32_t)C;
  float* slab0 = reinterpret_cast<float*>(ef_smem) + (size_t)(warp * EF_STAGES) * (slab_bytes / 4);
  uint64_t* bar0 = reinterpret_cast<uint64_t*>(ef_smem + (size_t)wpc * EF_STAGES * slab_bytes) + warp * EF_STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < EF_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ef_smem_u32(bar0 + s)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long n_tiles = p.tile_base[p.lv.num_levels];
  const long long gwarp = (long long)blockIdx.x * wpc + warp, gstride = (long long)gridDim.x * wpc;

  auto locate = [&](long long tile, int& l, long long& rec0, int& nrec) {
    l = 0;
#pragma unroll
    for (int k = 1; k < EF_MAX_LEVELS; ++k) if (k < p.lv.num_levels && tile >= p.tile_base[k]) l = k;
    rec0 = (tile - p.tile_base[l]) * 32;  // record index within the NB-image slice of level l
    const long long remain = (long long)p.NB * p.lv.anc_per_img[l] - rec0;
    nrec = remain < 32 ? (int)remain : 32;
  };
  auto issue = [&](long long tile, int s) {
    int l, nrec; long long rec0;
    locate(tile, l, rec0, nrec);
    const float* src = p.cls[l] + ((long long)p.B0 * p.lv.anc_per_img[l] + rec0) * C;
    const uint32_t bytes = (uint32_t)nrec * (uint32_t)C * 4u;
    float* dst = slab0 + (size_t)s * (slab_bytes / 4);
    if ((bytes & 15u) == 0u && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ef_smem_u32(bar0 + s)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ef_smem_u32(dst)),
                     "l"(src), "r"(bytes), "r"(ef_smem_u32(bar0 + s)) : "memory");
      }
    } else {
      for (int i = lane; i < nrec * C; i += 32) dst[i] = __ldg(src + i);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ef_smem_u32(bar0 + s)) : "memory");
    }
  };
  long long t_issue = gwarp;
#pragma unroll
  for (int s = 0; s < EF_STAGES; ++s) { if (t_issue < n_tiles) issue(t_issue, s); t_issue += gstride; }
  uint32_t phase = 0;
  int stage = 0;
  for (long long tile = gwarp; tile < n_tiles; tile += gstride) {
    {
      const uint32_t bar = ef_smem_u32(bar0 + stage);
      asm volatile("{\n.reg .pred p;\nEF_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra EF_DONE;\nbra EF_WAIT;\nEF_DONE:\n}\n" ::"r"(bar), "r"(phase) : "memory");
    }
    int l, nrec; long long rec0;
    locate(tile, l, rec0, nrec);
    const float* r = slab0 + (size_t)stage * (slab_bytes / 4) + lane * C;
    bool pass = false;
    int img = 0, cls = 0;
    float score = 0.f;
    uint32_t aidx = 0;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool fused_decode = p.rel[l] != nullptr;   // block-uniform
    int rin = 0, mi = 0;
    long long grec = 0;
    float m = 0.f;
    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < nrec) {
      const long long rec = rec0 + lane;
      const int api = p.lv.anc_per_img[l];
      const long long total = (long long)p.NB * api;
      if (total < 0x7fffffffLL) { img = (int)((uint32_t)rec / (uint32_t)api); rin = (int)((uint32_t)rec - (uint32_t)img * (uint32_t)api); }
      else { img = (int)(rec / api); rin = (int)(rec - (long long)img * api); }
      grec = (long long)(p.B0 + img) * api + rin;
      if (fused_decode) r4 = __ldcs(p.rel[l] + grec);   // issued before the class scan: coalesced, 16 bytes per lane
      // tf.math.argmax: first maximal index; reduce_max: the maximum (anc:172-174)
      m = r[0];
      for (int c = 1; c < C; ++c) { const float v = r[c]; if (v > m) { m = v; mi = c; } }
    }
    // the class scan was the last read of the slab: its refill is in flight during the decode and the append below
    __syncwarp();
    if (t_issue < n_tiles) issue(t_issue, stage);
    t_issue += gstride;
    if (lane < nrec) {
      if (fused_decode) {
        // convert_outputs_boxes (_boxes_decoder, anc:245-274) for every anchor: the dense decoded tensor is an output
        int y, x, an_i;
        ef_split(p.lv, l, rin, y, x, an_i);
        const AnchorBox an = ef_anchor(p.lv, l, y, x, an_i);
        const float yca = DM_DIV(DM_ADD(an.y2, an.y1), 2.0f), xca = DM_DIV(DM_ADD(an.x2, an.x1), 2.0f);
        const float ha = DM_SUB(an.y2, an.y1), wa = DM_SUB(an.x2, an.x1);
        const float w = DM_MUL(dm_expf(r4.w), wa), h = DM_MUL(dm_expf(r4.z), ha);
        const float yc = DM_ADD(DM_MUL(r4.x, ha), yca), xc = DM_ADD(DM_MUL(r4.y, wa), xca);
        const float hh = DM_DIV(h, 2.0f), hw = DM_DIV(w, 2.0f);
        box = make_float4(DM_SUB(yc, hh), DM_SUB(xc, hw), DM_ADD(yc, hh), DM_ADD(xc, hw));
        if (p.dec[l]) __stcs(p.dec[l] + grec, box);
      }
      if (mi != 0) {  // classes_mask = classes_id != 0 (anc:179)
        pass = true; cls = mi; score = m;
        aidx = (uint32_t)(p.lv.anchor_base[l] + rin);
        if (!fused_decode) box = __ldg(p.boxes[l] + grec);
      }
    }
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
        p.cand_box[slot] = box; p.cand_score[slot] = score; p.cand_cls[slot] = cls; p.cand_aidx[slot] = aidx;
        if (p.bitmap) atomicOr(&p.bitmap[(size_t)limg * p.bitmap_words + (aidx >> 5)], 1u << (aidx & 31u));
      }
      todo &= ~grp;
    }
    if (++stage == EF_STAGES) { stage = 0; phase ^= 1u; }
  }
}

struct EfFinalizeParams {
  int NB, n_img;
  NmsConfig cfg;
  const float4* cand_box; const float* cand_score; const int32_t* cand_cls; const uint32_t* cand_aidx;
  const int32_t* counts; const uint32_t* bitmap; int bitmap_words;
  int32_t* nms_pos;
  const unsigned long long* pre_keys; const uint32_t* pre_pos; const int* pre_count; const int* pre_elig;
  const unsigned long long* pre_khi;  // all nullptr when the pre-selection kernel was not run
  float* out_boxes; long long* out_cls; float* out_score; int32_t* out_sel_idx; int32_t* out_sel_anchor; int32_t* out_count;
};

template <int METRIC>
__global__ void __launch_bounds__(NMS_THREADS, 1) effdet_nms_finalize_kernel(EfFinalizeParams p) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  NMS_T(13);
  const int img = blockIdx.x;
  const size_t cbase = (size_t)img * p.n_img;
  NmsSegment seg;
  seg.boxes = reinterpret_cast<const float*>(p.cand_box + cbase);
  seg.scores = p.cand_score + cbase;
  seg.classes = nullptr;
  seg.order_id = p.cand_aidx + cbase;
  seg.n = p.counts[img];
  int32_t* pos = p.nms_pos + (size_t)img * p.cfg.max_out;
  NmsPre pre;
  if (p.pre_keys) {
    pre.keys = p.pre_keys + (size_t)img * NMS_PRE_CAP; pre.pos = p.pre_pos + (size_t)img * NMS_PRE_CAP;
    pre.count = p.pre_count + img; pre.eligible = p.pre_elig + img; pre.khi = p.pre_khi + img;
  }
  const int kept = nms_run_segment<METRIC, NmsLoadDirectAgnostic>(seg, p.cfg, pos, nms_smem, p.pre_keys ? &pre : nullptr);
  __syncthreads();
  NMS_T(11);
  if (threadIdx.x == 0) p.out_count[img] = kept;
  const size_t obase = (size_t)img * p.cfg.max_out;
  uint32_t* wprefix = reinterpret_cast<uint32_t*>(nms_smem);
  if (p.out_sel_idx && p.bitmap) {
    const uint32_t* bm = p.bitmap + (size_t)img * p.bitmap_words;
    if (threadIdx.x < 32) {
      uint32_t run = 0;
      for (int w0 = 0; w0 < p.bitmap_words; w0 += 32) {
        const int w = w0 + (int)threadIdx.x;
        const int c = (w < p.bitmap_words) ? __popc(bm[w]) : 0;
        const int inc = warp_scan_incl(c);
        if (w < p.bitmap_words) wprefix[w] = run + (uint32_t)(inc - c);
        run += (uint32_t)__shfl_sync(0xffffffffu, inc, 31);
      }
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < kept; k += blockDim.x) {
    const int q = pos[k];
    reinterpret_cast<float4*>(p.out_boxes)[obase + k] = p.cand_box[cbase + q];
    p.out_cls[obase + k] = (long long)p.cand_cls[cbase + q];               // tf.argmax -> int64 (anc:172)
    p.out_score[obase + k] = dm_sigmoidf(p.cand_score[cbase + q]);          // sigmoid only on survivors (anc:200)
    const uint32_t a = p.cand_aidx[cbase + q];
    if (p.out_sel_anchor) p.out_sel_anchor[obase + k] = (int32_t)a;
    if (p.out_sel_idx && p.bitmap) {
      const uint32_t wv = p.bitmap[(size_t)img * p.bitmap_words + (a >> 5)];
      p.out_sel_idx[obase + k] = (int32_t)(wprefix[a >> 5] + __popc(wv & ((1u << (a & 31u)) - 1u)));
    }
  }
  NMS_T(12);
}

#ifdef NMS_TRACE
extern "C" int b200_debug_nms_trace_effdet(long long* out_host64) {
  return cudaMemcpyFromSymbol(out_host64, g_nms_trace, sizeof(long long) * (64 + 4 * 32)) == cudaSuccess ? 0 : -1;
}
#endif

// ---- target assignment (anc:91-138) ------------------------------------------------------------------
struct EfAssignParams {
  EfLevels lv;
  int B, C;
  uint32_t magic_c;           // floor(2^32/C)+1 (exact quotient for dividends below 2^16 * C)
  float thr;
  const float* gt_boxes;      // [total,4] yxyx
  const int32_t* gt_classes;  // [total]
  const int32_t* gt_offsets;  // [B+1]
  float4* out_boxes[EF_MAX_LEVELS];        // (B,H,W,A,4)
  float* out_onehot[EF_MAX_LEVELS];        // (B,H,W,A,C), or null when out_class is used
  int32_t* out_class[EF_MAX_LEVELS];       // (B,H,W,A): class id instead of the one-hot row (sparse-target mode)
  unsigned char* out_mask[EF_MAX_LEVELS];  // (B,H,W,A,1)
  int cta_base[EF_MAX_LEVELS + 1];         // CTAs of 256 anchors, per (level, image)
  int chunks_per_img[EF_MAX_LEVELS];
};

#define EF_GT_TILE 128

__global__ void __launch_bounds__(256, 8) effdet_assign_kernel(EfAssignParams p) {
  __shared__ float4 s_gt[EF_GT_TILE];
  __shared__ float s_ga[EF_GT_TILE];
  int l = 0;
#pragma unroll
  for (int k = 1; k < EF_MAX_LEVELS; ++k) if (k < p.lv.num_levels && (int)blockIdx.x >= p.cta_base[k]) l = k;
  const int rc = blockIdx.x - p.cta_base[l];
  const int img = rc / p.chunks_per_img[l];
  const int chunk = rc - img * p.chunks_per_img[l];
  const int api = p.lv.anc_per_img[l];
  const int rin = chunk * 256 + (int)threadIdx.x;
  const bool active = rin < api;
  const int g_beg = p.gt_offsets[img], n_gt = p.gt_offsets[img + 1] - g_beg;
  BoxT an;
  an.c0 = an.c1 = an.c2 = an.c3 = an.area = an.at = 0.f;
  if (active) {
    int y, x, a;
    ef_split(p.lv, l, rin, y, x, a);
    const AnchorBox b = ef_anchor(p.lv, l, y, x, a);
    an = bm_prep(b.y1, b.x1, b.y2, b.x2, B200_METRIC_EFF_IOU);
  }
  // Bounding box of the CTA's 256 anchors (a run of cells of one row): a GT that does not overlap it has IoU exactly 0
  // with every anchor here and, for thr > 0, can neither match nor change the outcome (an all-zero row stays
  // unmatched whatever its argmax), so each GT tile is culled once per CTA and the anchors walk only the survivors —
  // in ascending GT order, which keeps tf.argmax's first-maximal-index rule.
  __shared__ float s_bb[8][4];
  __shared__ uint32_t s_mask[EF_GT_TILE / 32];
  {
    float y1 = active ? an.c0 : INFINITY, x1 = active ? an.c1 : INFINITY;
    float y2 = active ? an.c2 : -INFINITY, x2 = active ? an.c3 : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      y1 = fminf(y1, __shfl_xor_sync(0xffffffffu, y1, o)); x1 = fminf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
      y2 = fmaxf(y2, __shfl_xor_sync(0xffffffffu, y2, o)); x2 = fmaxf(x2, __shfl_xor_sync(0xffffffffu, x2, o));
    }
    if ((threadIdx.x & 31) == 0) { float* d = s_bb[threadIdx.x >> 5]; d[0] = y1; d[1] = x1; d[2] = y2; d[3] = x2; }
  }
  __syncthreads();
  float by1 = s_bb[0][0], bx1 = s_bb[0][1], by2 = s_bb[0][2], bx2 = s_bb[0][3];
#pragma unroll
  for (int w = 1; w < 8; ++w) {
    by1 = fminf(by1, s_bb[w][0]); bx1 = fminf(bx1, s_bb[w][1]); by2 = fmaxf(by2, s_bb[w][2]); bx2 = fmaxf(bx2, s_bb[w][3]);
  }
  const bool cull = p.thr > 0.0f;
  float best = -INFINITY;
  int best_i = 0;
  for (int g0 = 0; g0 < n_gt; g0 += EF_GT_TILE) {
    const int m = min(EF_GT_TILE, n_gt - g0);
    __syncthreads();
    if (threadIdx.x < EF_GT_TILE) {
      bool relevant = false;
      if ((int)threadIdx.x < m) {
        const float* q = p.gt_boxes + 4 * (size_t)(g_beg + g0 + threadIdx.x);
        const BoxT g = bm_prep(q[0], q[1], q[2], q[3], B200_METRIC_EFF_IOU);
        s_gt[threadIdx.x] = make_float4(g.c0, g.c1, g.c2, g.c3);
        s_ga[threadIdx.x] = g.area;
        // positive overlap with the bounding box in both axes (false for NaN coordinates, which can never be chosen)
        relevant = !cull || ((DM_SUB(dm_min(by2, g.c2), dm_max(by1, g.c0)) > 0.0f) && (DM_SUB(dm_min(bx2, g.c3), dm_max(bx1, g.c1)) > 0.0f));
      }
      const uint32_t bits = __ballot_sync(0xffffffffu, relevant);
      if ((threadIdx.x & 31) == 0) s_mask[threadIdx.x >> 5] = bits;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int w = 0; w < EF_GT_TILE / 32; ++w) {
        uint32_t bits = s_mask[w];
        while (bits) {
          const int g = (w << 5) + __ffs(bits) - 1;
          bits &= bits - 1u;
          const float4 c = s_gt[g];
          // clamped intersection extents (eiou:58-64); zero overlap -> iou = divide_no_nan(0, union) = +0 exactly
          const float iw = dm_max(0.0f, DM_SUB(dm_min(an.c3, c.w), dm_max(an.c1, c.y)));
          const float ih = dm_max(0.0f, DM_SUB(dm_min(an.c2, c.z), dm_max(an.c0, c.x)));
          float v = 0.0f;
          if (iw > 0.0f && ih > 0.0f) {
            const float inter = DM_MUL(iw, ih);
            v = bm_dnn(inter, DM_SUB(DM_ADD(an.area, s_ga[g]), inter));
          } else if (!(iw == iw) || !(ih == ih)) {
            BoxT gb; gb.c0 = c.x; gb.c1 = c.y; gb.c2 = c.z; gb.c3 = c.w; gb.area = s_ga[g]; gb.at = 0.f;
            v = bm_metric(an, gb, B200_METRIC_EFF_IOU);
          }
          if (v > best) { best = v; best_i = g0 + g; }  // tf.argmax: first maximal index
        }
      }
    }
  }
  // encode + outputs
  const bool matched = active && (n_gt > 0) && (best >= p.thr);  // iou_max >= thr (anc:121)
  int cls = 0;
  float4 enc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (matched) {
    const float* q = p.gt_boxes + 4 * (size_t)(g_beg + best_i);
    const float by1 = q[0], bx1 = q[1], by2 = q[2], bx2 = q[3];
    cls = p.gt_classes[g_beg + best_i];
    // _boxes_encoder, anc:231-243
    const float yca = DM_DIV(DM_ADD(an.c2, an.c0), 2.0f), xca = DM_DIV(DM_ADD(an.c3, an.c1), 2.0f);
    const float ha = dm_max(1e-8f, DM_SUB(an.c2, an.c0)), wa = dm_max(1e-8f, DM_SUB(an.c3, an.c1));
    const float yc = DM_DIV(DM_ADD(by2, by1), 2.0f), xc = DM_DIV(DM_ADD(bx2, bx1), 2.0f);
    const float h = dm_max(1e-8f, DM_SUB(by2, by1)), w = dm_max(1e-8f, DM_SUB(bx2, bx1));
    enc.x = DM_DIV(DM_SUB(yc, yca), ha);
    enc.y = DM_DIV(DM_SUB(xc, xca), wa);
    enc.z = dm_logf(DM_DIV(h, ha));
    enc.w = dm_logf(DM_DIV(w, wa));
  }
  const size_t abase = (size_t)img * api;
  if (active) {
    p.out_boxes[l][abase + rin] = enc;
    p.out_mask[l][abase + rin] = matched ? 1 : 0;
  }
  if (p.out_class[l]) {  // sparse-target mode: the class id stands for the one-hot row (block-uniform branch)
    if (active) p.out_class[l][abase + rin] = cls;
    return;
  }
  // one-hot rows (C floats per anchor, class 0 = background for unmatched anchors, anc:131-133; tf.one_hot: out of
  // range -> zeros): the CTA's 256 rows are one contiguous span of rows*C floats, written once with aligned 16-byte
  // streaming stores; the 1.0 of row r sits at flat element r*C + cls[r] of the span.
  __shared__ int s_hot[256];
  s_hot[threadIdx.x] = (active && cls >= 0 && cls < p.C) ? (int)threadIdx.x * p.C + cls : -1;
  __syncthreads();
  const int rows = min(256, api - chunk * 256);
  float* dst = p.out_onehot[l] + (abase + (size_t)chunk * 256) * p.C;
  const int n_el = rows * p.C;
  const int head = min(n_el, (int)(((16u - ((unsigned)(uintptr_t)dst & 15u)) & 15u) >> 2));
  const int n_vec = p.C >= 3 ? (n_el - head) >> 2 : 0;  // rows narrower than 3 take the scalar path below
  float4* dst4 = reinterpret_cast<float4*>(dst + head);
  for (int i = threadIdx.x; i < n_vec; i += 256) {
    const int e = head + 4 * i;                                   // first flat element of this float4
    const int r = (int)__umulhi((uint32_t)e, p.magic_c);          // e / C (magic_c = floor(2^32/C)+1, e < 2^16 * C)
    const int h0 = s_hot[r], h1 = (r + 1 < rows) ? s_hot[r + 1] : -1;  // a float4 touches at most rows r and r+1 when C >= 3
    float4 v;
    v.x = (e == h0 || e == h1) ? 1.0f : 0.0f;
    v.y = (e + 1 == h0 || e + 1 == h1) ? 1.0f : 0.0f;
    v.z = (e + 2 == h0 || e + 2 == h1) ? 1.0f : 0.0f;
    v.w = (e + 3 == h0 || e + 3 == h1) ? 1.0f : 0.0f;
    __stcs(dst4 + i, v);
  }
  // unaligned head / tail elements (and every element of rows narrower than 3): scalar stores
  for (int e = threadIdx.x; e < n_el; e += 256) {
    if (e >= head && e < head + (n_vec << 2)) continue;
    const int r = e / p.C;
    dst[e] = (s_hot[r] == e) ? 1.0f : 0.0f;
  }
}

// ---- host side ------------------------------------------------------------------------------------
static int ef_fill_levels(EfLevels& lv, int num_levels, const int32_t* hw, int A, const float* table_dev) {
  if (num_levels < 1 || num_levels > EF_MAX_LEVELS || A < 1) return -1;
  lv.num_levels = num_levels; lv.A = A; lv.table = table_dev;
  int base = 0, off = 0;
  for (int l = 0; l < EF_MAX_LEVELS; ++l) {
    if (l < num_levels) {
      lv.h[l] = hw[2 * l]; lv.w[l] = hw[2 * l + 1];
      if (lv.h[l] <= 0 || lv.w[l] <= 0) return -1;
      lv.anc_per_img[l] = lv.h[l] * lv.w[l] * A;
      lv.anchor_base[l] = base; base += lv.anc_per_img[l];
      lv.tab_off[l] = off; off += lv.h[l] + lv.w[l] + 2 * A;
    } else {
      lv.h[l] = lv.w[l] = 1; lv.anc_per_img[l] = 0; lv.anchor_base[l] = base; lv.tab_off[l] = off;
    }
  }
  return base;
}

extern "C" size_t b200_effdet_table_floats(int num_levels, const int32_t* hw, int A) {
  size_t n = 0;
  for (int l = 0; l < num_levels; ++l) n += (size_t)hw[2 * l] + hw[2 * l + 1] + 2 * (size_t)A;
  return n;
}

extern "C" int b200_effdet_anchors(int num_levels, const int32_t* hw, int A, const float* table_dev, int level,
                                   float* out, void* stream) {
  EfLevels lv;
  B200_REQUIRE(hw && table_dev && out, B200_ERR_BAD_ARG, "b200_effdet_anchors: null argument");
  B200_REQUIRE(ef_fill_levels(lv, num_levels, hw, A, table_dev) >= 0, B200_ERR_BAD_ARG, "b200_effdet_anchors: bad level spec");
  B200_REQUIRE(level >= 0 && level < num_levels, B200_ERR_BAD_ARG, "b200_effdet_anchors: bad level %d", level);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, B200_ERR_BAD_ARG, "b200_effdet_anchors: out not 16-byte aligned");
  const int n = lv.anc_per_img[level];
  effdet_anchors_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(lv, level, reinterpret_cast<float4*>(out));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_effdet_decode(int num_levels, const int32_t* hw, int A, const float* table_dev, int B,
                                  const float* const rel[], float* const out[], void* stream) {
  EfDecodeParams p;
  B200_REQUIRE(hw && table_dev && rel && out, B200_ERR_BAD_ARG, "b200_effdet_decode: null argument");
  B200_REQUIRE(ef_fill_levels(p.lv, num_levels, hw, A, table_dev) >= 0, B200_ERR_BAD_ARG, "b200_effdet_decode: bad level spec");
  B200_REQUIRE(B >= 0, B200_ERR_BAD_ARG, "b200_effdet_decode: negative batch");
  if (B == 0) return B200_OK;
  p.B = B;
  long long cum = 0;
  for (int l = 0; l < EF_MAX_LEVELS; ++l) {
    p.elem_base[l] = cum;
    if (l < num_levels) {
      B200_REQUIRE(rel[l] && out[l], B200_ERR_BAD_ARG, "b200_effdet_decode: null level %d", l);
      B200_REQUIRE(((reinterpret_cast<uintptr_t>(rel[l]) | reinterpret_cast<uintptr_t>(out[l])) & 15) == 0, B200_ERR_BAD_ARG,
                   "b200_effdet_decode: level %d not 16-byte aligned", l);
      p.rel[l] = reinterpret_cast<const float4*>(rel[l]);
      p.out[l] = reinterpret_cast<float4*>(out[l]);
      cum += (long long)B * p.lv.anc_per_img[l];
    } else { p.rel[l] = nullptr; p.out[l] = nullptr; }
  }
  p.elem_base[EF_MAX_LEVELS] = cum;
  for (int l = num_levels; l <= EF_MAX_LEVELS; ++l) p.elem_base[l] = cum;
  long long blocks = (cum + 255) / 256;
  const long long cap = (long long)b200_sm_count() * 64;  // many short CTAs: see el_cta_plan
  if (blocks > cap) blocks = cap;
  effdet_decode_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

struct EfWs { size_t counts, pre_count, pre_elig, bitmap, box, score, cls, aidx, pos, pre_khi, pre_keys, pre_pos, total; int bitmap_words; };
static EfWs ef_ws_layout(int NB, int n_img, int max_out) {
  EfWs w;
  size_t o = 0;
  w.bitmap_words = (n_img + 31) / 32;
  w.counts = o; o = b200_align_up(o + sizeof(int32_t) * NB, 256);
  w.pre_count = o; o = b200_align_up(o + sizeof(int32_t) * NB, 256);   // zeroed together with counts
  w.pre_elig = o; o = b200_align_up(o + sizeof(int32_t) * NB, 256);
  w.bitmap = o; o = b200_align_up(o + sizeof(uint32_t) * (size_t)NB * w.bitmap_words, 256);
  w.box = o; o = b200_align_up(o + sizeof(float4) * (size_t)NB * n_img, 256);
  w.score = o; o = b200_align_up(o + sizeof(float) * (size_t)NB * n_img, 256);
  w.cls = o; o = b200_align_up(o + sizeof(int32_t) * (size_t)NB * n_img, 256);
  w.aidx = o; o = b200_align_up(o + sizeof(uint32_t) * (size_t)NB * n_img, 256);
  w.pos = o; o = b200_align_up(o + sizeof(int32_t) * (size_t)NB * max_out, 256);
  w.pre_khi = o; o = b200_align_up(o + sizeof(unsigned long long) * NB, 256);
  w.pre_keys = o; o = b200_align_up(o + sizeof(unsigned long long) * (size_t)NB * NMS_PRE_CAP, 256);
  w.pre_pos = o; o = b200_align_up(o + sizeof(uint32_t) * (size_t)NB * NMS_PRE_CAP, 256);
  w.total = o;
  return w;
}

__global__ void __launch_bounds__(NMS_THREADS, 1) effdet_nms_pivot_kernel(NmsPreselectParams p) {
  extern __shared__ __align__(16) unsigned char pre_smem[];
  nms_pivot_body(p, pre_smem);
}
__global__ void __launch_bounds__(256) effdet_nms_pregather_kernel(NmsPreselectParams p) { nms_pregather_body(p); }

// The NMS tail shared by b200_effdet_postprocess and the fused entry points: (large segments) pivot + multi-CTA
// pre-gather of the first window, then one CTA per image.
static int ef_launch_nms(const EfFilterParams& fp, const EfWs& ws, unsigned char* wsb, int num_images, int n_img, int max_out,
                         float iou_thr, float score_thr, int metric, float* out_boxes, long long* out_class_id, float* out_score,
                         int32_t* out_sel_idx, int32_t* out_sel_anchor, int32_t* out_count, cudaStream_t stream) {
  EfFinalizeParams np;
  np.NB = num_images; np.n_img = n_img;
  np.cfg.metric = metric; np.cfg.mode = B200_NMS_AGNOSTIC; np.cfg.iou_thr = iou_thr; np.cfg.score_thr = score_thr;
  np.cfg.use_score_thr = 1; np.cfg.max_out = max_out;
  np.cand_box = fp.cand_box; np.cand_score = fp.cand_score; np.cand_cls = fp.cand_cls; np.cand_aidx = fp.cand_aidx;
  np.counts = fp.counts; np.bitmap = fp.bitmap; np.bitmap_words = ws.bitmap_words;
  np.nms_pos = reinterpret_cast<int32_t*>(wsb + ws.pos);
  np.pre_keys = nullptr; np.pre_pos = nullptr; np.pre_count = nullptr; np.pre_elig = nullptr; np.pre_khi = nullptr;
  if (n_img > NMS_PRE_MIN_N) {
    // large segments: several CTAs per image gather the first NMS window so the single NMS CTA does not have to
    // scan hundreds of thousands of scores alone
    NmsPreselectParams pp;
    int slices = (5 * b200_sm_count() + num_images - 1) / num_images;   // ~5 CTAs of 256 threads (46 registers) per SM: one wave
    if (slices < 1) slices = 1;
    if (slices > 64) slices = 64;
    pp.scores = fp.cand_score; pp.order_id = fp.cand_aidx; pp.counts = fp.counts; pp.stride = n_img; pp.slices = slices;
    pp.use_score_thr = 1; pp.score_thr = score_thr;
    pp.keys = reinterpret_cast<unsigned long long*>(wsb + ws.pre_keys);
    pp.pos = reinterpret_cast<uint32_t*>(wsb + ws.pre_pos);
    pp.count = reinterpret_cast<int*>(wsb + ws.pre_count);
    pp.eligible = reinterpret_cast<int*>(wsb + ws.pre_elig);
    pp.khi = reinterpret_cast<unsigned long long*>(wsb + ws.pre_khi);
    B200_CUDA(cudaFuncSetAttribute(effdet_nms_pivot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NMS_PIVOT_SMEM));
    effdet_nms_pivot_kernel<<<num_images, NMS_THREADS, NMS_PIVOT_SMEM, stream>>>(pp);
    B200_LAUNCH_CHECK();
    effdet_nms_pregather_kernel<<<num_images * slices, 256, 0, stream>>>(pp);
    B200_LAUNCH_CHECK();
    np.pre_keys = pp.keys; np.pre_pos = pp.pos; np.pre_count = pp.count; np.pre_elig = pp.eligible; np.pre_khi = pp.khi;
  }
  np.out_boxes = out_boxes; np.out_cls = out_class_id; np.out_score = out_score; np.out_sel_idx = out_sel_idx;
  np.out_sel_anchor = out_sel_anchor; np.out_count = out_count;
  size_t smem2 = nms_smem_bytes(max_out);
  if (smem2 < (size_t)ws.bitmap_words * 4) smem2 = (size_t)ws.bitmap_words * 4;
  B200_REQUIRE(smem2 <= 220 * 1024, B200_ERR_UNSUPPORTED, "b200_effdet_postprocess: too many anchors per image for the rank table");
#define EF_LAUNCH(M)                                                                                                     \
  B200_CUDA(cudaFuncSetAttribute(effdet_nms_finalize_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
  effdet_nms_finalize_kernel<M><<<num_images, NMS_THREADS, smem2, stream>>>(np)
  NMS_DISPATCH_METRIC(metric, EF_LAUNCH)
#undef EF_LAUNCH
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" size_t b200_effdet_postprocess_workspace_bytes(int num_levels, const int32_t* hw, int A, int num_images, int max_out) {
  int n_img = 0;
  for (int l = 0; l < num_levels; ++l) n_img += hw[2 * l] * hw[2 * l + 1] * A;
  return ef_ws_layout(num_images, n_img, max_out).total;
}

// boxes: decoded boxes (the drop-in convert_outputs_one), or null with rel / table_dev given: decode inside the filter
// pass (convert_outputs_boxes + convert_outputs_one in one sweep; dec_out[l] receives the dense decoded tensor)
static int ef_post_impl(int num_levels, const int32_t* hw, int A, const float* table_dev, int C, int B, int first_image,
                        int num_images, const float* const boxes[], const float* const rel[], float* const dec_out[],
                        const float* const classes[], int max_out, float iou_thr, float score_thr, int metric, float* out_boxes,
                        long long* out_class_id, float* out_score, int32_t* out_sel_idx,
                        int32_t* out_sel_anchor, int32_t* out_count, void* workspace,
                        size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EfFilterParams fp;
  B200_REQUIRE(hw && (boxes || (rel && table_dev)) && classes, B200_ERR_BAD_ARG, "b200_effdet_postprocess: null argument");
  const int n_img = ef_fill_levels(fp.lv, num_levels, hw, A, table_dev);
  B200_REQUIRE(n_img >= 0, B200_ERR_BAD_ARG, "b200_effdet_postprocess: bad level spec");
  B200_REQUIRE(C >= 1 && C <= 400, B200_ERR_UNSUPPORTED, "b200_effdet_postprocess: classes_num %d outside [1,400]", C);
  B200_REQUIRE(first_image >= 0 && num_images >= 0 && first_image + num_images <= B, B200_ERR_BAD_ARG, "b200_effdet_postprocess: image range outside the batch");
  B200_REQUIRE(metric >= B200_METRIC_EFF_IOU && metric <= B200_METRIC_EFF_CIOU, B200_ERR_BAD_ARG, "b200_effdet_postprocess: iou_type must be iou/giou/diou/ciou");
  B200_REQUIRE(max_out >= 1 && max_out <= NMS_MAX_OUT_LIMIT, B200_ERR_UNSUPPORTED, "b200_effdet_postprocess: max_out %d outside [1,%d]", max_out, NMS_MAX_OUT_LIMIT);
  if (num_images == 0) return B200_OK;
  B200_REQUIRE(out_boxes && out_class_id && out_score && out_count, B200_ERR_BAD_ARG, "b200_effdet_postprocess: null output");
  EfWs ws = ef_ws_layout(num_images, n_img, max_out);
  B200_REQUIRE(workspace && workspace_bytes >= ws.total, B200_ERR_WORKSPACE, "b200_effdet_postprocess: workspace %zu < required %zu", workspace_bytes, ws.total);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, B200_ERR_BAD_ARG, "b200_effdet_postprocess: workspace not 256-byte aligned");
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  fp.B0 = first_image; fp.NB = num_images; fp.C = C; fp.B = B; fp.n_img = n_img;
  long long tb = 0;
  for (int l = 0; l < EF_MAX_LEVELS; ++l) {
    fp.tile_base[l] = tb;
    if (l < num_levels) {
      const float* bl = boxes ? boxes[l] : rel[l];
      B200_REQUIRE(bl && classes[l], B200_ERR_BAD_ARG, "b200_effdet_postprocess: null level %d", l);
      B200_REQUIRE(((reinterpret_cast<uintptr_t>(bl) | (dec_out && dec_out[l] ? reinterpret_cast<uintptr_t>(dec_out[l]) : 0)) & 15) == 0, B200_ERR_BAD_ARG,
                   "b200_effdet_postprocess: boxes level %d not 16-byte aligned", l);
      fp.cls[l] = classes[l];
      fp.boxes[l] = boxes ? reinterpret_cast<const float4*>(boxes[l]) : nullptr;
      fp.rel[l] = boxes ? nullptr : reinterpret_cast<const float4*>(rel[l]);
      fp.dec[l] = (!boxes && dec_out) ? reinterpret_cast<float4*>(dec_out[l]) : nullptr;
      tb += ((long long)num_images * fp.lv.anc_per_img[l] + 31) / 32;
    } else { fp.cls[l] = nullptr; fp.boxes[l] = nullptr; fp.rel[l] = nullptr; fp.dec[l] = nullptr; }
  }
  for (int l = num_levels; l <= EF_MAX_LEVELS; ++l) fp.tile_base[l] = tb;
  fp.cand_box = reinterpret_cast<float4*>(wsb + ws.box);
  fp.cand_score = reinterpret_cast<float*>(wsb + ws.score);
  fp.cand_cls = reinterpret_cast<int32_t*>(wsb + ws.cls);
  fp.cand_aidx = reinterpret_cast<uint32_t*>(wsb + ws.aidx);
  fp.counts = reinterpret_cast<int32_t*>(wsb + ws.counts);
  fp.bitmap = out_sel_idx ? reinterpret_cast<uint32_t*>(wsb + ws.bitmap) : nullptr;
  fp.bitmap_words = ws.bitmap_words;
  B200_CUDA(cudaMemsetAsync(wsb, 0, out_sel_idx ? ws.box : ws.bitmap, stream));
  const uint32_t slab = 128u * (uint32_t)C;
  int warps = 8;
  while (warps > 1 && (size_t)warps * EF_STAGES * slab + 256 > 200 * 1024) warps >>= 1;
  const size_t smem1 = (size_t)warps * EF_STAGES * slab + sizeof(uint64_t) * warps * EF_STAGES + 16;
  B200_CUDA(cudaFuncSetAttribute(effdet_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  long long want = (tb + warps - 1) / warps;
  int grid = (int)(want < (long long)b200_sm_count() ? want : (long long)b200_sm_count());
  if (grid < 1) grid = 1;
  effdet_filter_kernel<<<grid, warps * 32, smem1, stream>>>(fp);
  B200_LAUNCH_CHECK();

  return ef_launch_nms(fp, ws, wsb, num_images, n_img, max_out, iou_thr, score_thr, metric, out_boxes, out_class_id, out_score,
                       out_sel_idx, out_sel_anchor, out_count, stream);
}

extern "C" int b200_effdet_postprocess(int num_levels, const int32_t* hw, int A, int C, int B, int first_image,
                                       int num_images, const float* const boxes[], const float* const classes[],
                                       int max_out, float iou_thr, float score_thr, int metric, float* out_boxes,
                                       long long* out_class_id, float* out_score, int32_t* out_sel_idx,
                                       int32_t* out_sel_anchor, int32_t* out_count, void* workspace,
                                       size_t workspace_bytes, void* stream_) {
  B200_REQUIRE(boxes, B200_ERR_BAD_ARG, "b200_effdet_postprocess: null argument");
  return ef_post_impl(num_levels, hw, A, nullptr, C, B, first_image, num_images, boxes, nullptr, nullptr, classes, max_out, iou_thr,
                      score_thr, metric, out_boxes, out_class_id, out_score, out_sel_idx, out_sel_anchor, out_count, workspace,
                      workspace_bytes, stream_);
}

static int ef_assign_impl(int num_levels, const int32_t* hw, int A, const float* table_dev, int C, int B,
                          const float* gt_boxes, const int32_t* gt_classes, const int32_t* gt_offsets, float iou_thr,
                          float* const out_boxes[], float* const out_onehot[], int32_t* const out_class[],
                          unsigned char* const out_mask[], void* stream) {
  EfAssignParams p;
  B200_REQUIRE(hw && table_dev && gt_offsets && out_boxes && (out_onehot || out_class) && out_mask, B200_ERR_BAD_ARG, "b200_effdet_assign_targets: null argument");
  B200_REQUIRE(ef_fill_levels(p.lv, num_levels, hw, A, table_dev) >= 0, B200_ERR_BAD_ARG, "b200_effdet_assign_targets: bad level spec");
  B200_REQUIRE(B >= 0 && C >= 1, B200_ERR_BAD_ARG, "b200_effdet_assign_targets: bad sizes");
  if (B == 0) return B200_OK;
  p.B = B; p.C = C; p.thr = iou_thr;
  p.magic_c = C == 1 ? 0u : (uint32_t)((1ull << 32) / (unsigned long long)C + 1ull);
  p.gt_boxes = gt_boxes; p.gt_classes = gt_classes; p.gt_offsets = gt_offsets;
  int cta = 0;
  for (int l = 0; l < EF_MAX_LEVELS; ++l) {
    p.cta_base[l] = cta;
    if (l < num_levels) {
      B200_REQUIRE(out_boxes[l] && (out_onehot ? out_onehot[l] != nullptr : out_class[l] != nullptr) && out_mask[l], B200_ERR_BAD_ARG,
                   "b200_effdet_assign_targets: null level %d", l);
      B200_REQUIRE((reinterpret_cast<uintptr_t>(out_boxes[l]) & 15) == 0, B200_ERR_BAD_ARG, "b200_effdet_assign_targets: out_boxes level %d not 16-byte aligned", l);
      p.out_boxes[l] = reinterpret_cast<float4*>(out_boxes[l]);
      p.out_onehot[l] = out_onehot ? out_onehot[l] : nullptr;
      p.out_class[l] = out_onehot ? nullptr : out_class[l];
      p.out_mask[l] = out_mask[l];
      p.chunks_per_img[l] = (p.lv.anc_per_img[l] + 255) / 256;
      cta += p.chunks_per_img[l] * B;
    } else { p.out_boxes[l] = nullptr; p.out_onehot[l] = nullptr; p.out_class[l] = nullptr; p.out_mask[l] = nullptr; p.chunks_per_img[l] = 1; }
  }
  for (int l = num_levels; l <= EF_MAX_LEVELS; ++l) p.cta_base[l] = cta;
  effdet_assign_kernel<<<cta, 256, 0, (cudaStream_t)stream>>>(p);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_effdet_assign_targets(int num_levels, const int32_t* hw, int A, const float* table_dev, int C, int B,
                                          const float* gt_boxes, const int32_t* gt_classes, const int32_t* gt_offsets,
                                          float iou_thr, float* const out_boxes[], float* const out_onehot[],
                                          unsigned char* const out_mask[], void* stream) {
  B200_REQUIRE(out_onehot, B200_ERR_BAD_ARG, "b200_effdet_assign_targets: null argument");
  return ef_assign_impl(num_levels, hw, A, table_dev, C, B, gt_boxes, gt_classes, gt_offsets, iou_thr, out_boxes, out_onehot,
                        nullptr, out_mask, stream);
}

extern "C" int b200_effdet_assign_targets_indexed(int num_levels, const int32_t* hw, int A, const float* table_dev, int C, int B,
                                                  const float* gt_boxes, const int32_t* gt_classes, const int32_t* gt_offsets,
                                                  float iou_thr, float* const out_boxes[], int32_t* const out_class[],
                                                  unsigned char* const out_mask[], void* stream) {
  B200_REQUIRE(out_class, B200_ERR_BAD_ARG, "b200_effdet_assign_targets_indexed: null argument");
  return ef_assign_impl(num_levels, hw, A, table_dev, C, B, gt_boxes, gt_classes, gt_offsets, iou_thr, out_boxes, nullptr,
                        out_class, out_mask, stream);
}

// ---- fused eval-step stream (test_step, efficientnet/efficientdet_net_train.py:135-169) ----------------------------
// The reference's test_step reads the class logits twice: FocalLoss (:150) and, after convert_outputs_boxes (:153), the
// per-image argmax / max of convert_outputs_one (anchors.py:172-174).  Here ONE pass over every level does
//   * focal loss of the tile's logits against the one-hot targets (flat 128-bit streaming loads, both tensors),
//   * argmax / max over each anchor's C logits (the tile also goes to shared memory; four lanes per anchor),
//   * box decode from the head offsets and the anchor table (anchors.py:245-274) -> the dense decoded tensor,
//   * Huber box loss + positive count (box_loss.py:17-29), and
//   * the append of every non-background anchor to its image's candidate list (anchors.py:179-189).
// (Without targets — convert_outputs_boxes + convert_outputs_one alone — the per-warp bulk-copy filter above is the faster
// stream and takes the decode along: b200_effdet_decode_postprocess.)
#define EFU_TILE 64      // anchors per tile
#define EFU_THREADS 256
#ifndef EFU_INFLIGHT
#define EFU_INFLIGHT 4   // float4 of each class tensor a stream thread has in flight
#endif

struct EfFusedParams {
  EfLevels lv;
  int B, C, n_img;
  const float* cls[EF_MAX_LEVELS];        // (B,H,W,A,C) logits
  const float* cls_true[EF_MAX_LEVELS];   // (B,H,W,A,C) targets (WITH_LOSS)
  const float4* rel[EF_MAX_LEVELS];       // (B,H,W,A,4) head offsets ty,tx,th,tw
  const float4* box_true[EF_MAX_LEVELS];  // (WITH_LOSS)
  const unsigned char* mask[EF_MAX_LEVELS];
  float4* dec[EF_MAX_LEVELS];             // decoded boxes out
  long long tile_base[EF_MAX_LEVELS + 1];
  float4* cand_box; float* cand_score; int32_t* cand_cls; uint32_t* cand_aidx; int32_t* counts;
  uint32_t* bitmap; int bitmap_words;
  float alpha, gamma, delta, label_smoothing;
  double* partials;                       // [gridDim.x][num_levels][3] focal, huber, positives
};

// Warp roles inside a CTA (tile = 64 anchors):
//   warps 2-7 (CLASS stream, 192 threads): the tile's logits (and one-hot targets) arrive in shared memory by 1-D bulk
//     asynchronous copies (cp.async.bulk + mbarrier, two stages: the copy of the CTA's next tile is in flight while this
//     one is processed); focal terms from flat 128-bit shared-memory reads, then the per-anchor argmax with three lanes
//     per anchor (partial (max, first index) per third of the C logits -> shared memory);
//   warps 0-1 (BOX half, lane <-> anchor): head offsets / box targets / mask straight from global memory, decode, Huber,
//     and — behind the tile's single CTA barrier — the fold of the three argmax partials and the candidate append.
// Both halves are ~600-800 instructions per warp and tile and independent, so they run side by side.
#define EFU_BOX_WARPS 2
#define EFU_STREAM_THREADS (EFU_THREADS - 32 * EFU_BOX_WARPS)
#define EFU_STAGES 2

template <bool WITH_LOSS, bool G15>
__global__ void __launch_bounds__(EFU_THREADS, WITH_LOSS ? 2 : 4) effdet_stream_kernel(EfFusedParams p) {
  extern __shared__ __align__(128) unsigned char efu_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C;
  const uint32_t tile_bytes = (uint32_t)EFU_TILE * (uint32_t)C * 4u;
  // stage s: [x tile][y tile (WITH_LOSS)]
  const uint32_t stage_bytes = tile_bytes * (WITH_LOSS ? 2u : 1u);
  __shared__ __align__(8) uint64_t s_bar[EFU_STAGES];
  __shared__ double s_part[EF_MAX_LEVELS][3];
  __shared__ double s_red[EFU_THREADS / 32][3];
  __shared__ float s_m[2][EFU_TILE][3];
  __shared__ int s_mi[2][EFU_TILE][3];
  if (tid < EF_MAX_LEVELS * 3) s_part[tid / 3][tid % 3] = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < EFU_STAGES; ++s) bc_mbar_init(&s_bar[s], 1);
    bc_fence_mbar_init();
  }
  __syncthreads();
  const long long n_tiles = p.tile_base[p.lv.num_levels];
  double focal = 0.0;
  float hub = 0.f;
  unsigned int pos = 0;
  int cur_level = -1;
  auto flush = [&](int level) {   // block-reduce this thread's running sums into the level's slot
    double v0 = warp_sum_d(focal), v1 = warp_sum_d((double)hub), v2 = warp_sum_d((double)pos);
    if (lane == 0) { s_red[warp][0] = v0; s_red[warp][1] = v1; s_red[warp][2] = v2; }
    __syncthreads();
    if (tid < 3) {
      double sacc = 0.0;
      for (int w = 0; w < EFU_THREADS / 32; ++w) sacc += s_red[w][tid];
      s_part[level][tid] += sacc;
    }
    __syncthreads();
    focal = 0.0; hub = 0.f; pos = 0;
  };
  // (level, first anchor, anchors, bulk-copyable) of a tile
  auto locate = [&](long long t, int& l, long long& rec0, int& nrec) {
    l = 0;
#pragma unroll
    for (int k = 1; k < EF_MAX_LEVELS; ++k) if (k < p.lv.num_levels && t >= p.tile_base[k]) l = k;
    rec0 = (t - p.tile_base[l]) * EFU_TILE;
    const long long remain = (long long)p.B * p.lv.anc_per_img[l] - rec0;
    nrec = remain < EFU_TILE ? (int)remain : EFU_TILE;
  };
  auto bulk_ok = [&](int l, long long rec0, int nrec) {
    const uintptr_t ax = reinterpret_cast<uintptr_t>(p.cls[l] + rec0 * C);
    const uintptr_t ay = WITH_LOSS ? reinterpret_cast<uintptr_t>(p.cls_true[l] + rec0 * C) : 0;
    return (((ax | ay) & 15) == 0) && ((((uint32_t)nrec * (uint32_t)C * 4u) & 15u) == 0u);
  };
  auto issue = [&](long long t, int s) {   // one thread: start the copies of tile t into stage s
    int l, nrec; long long rec0;
    locate(t, l, rec0, nrec);
    if (!bulk_ok(l, rec0, nrec)) return;   // ragged / unaligned tile: read straight from global memory when it is consumed
    const uint32_t bytes = (uint32_t)nrec * (uint32_t)C * 4u;
    unsigned char* dst = efu_smem + (size_t)s * stage_bytes;
    bc_fence_proxy_async();
    bc_mbar_expect_tx(&s_bar[s], bytes * (WITH_LOSS ? 2u : 1u));
    bc_bulk_g2s(dst, p.cls[l] + rec0 * C, bytes, &s_bar[s]);
    if (WITH_LOSS) bc_bulk_g2s(dst + tile_bytes, p.cls_true[l] + rec0 * C, bytes, &s_bar[s]);
  };
  const bool box_warp = warp < EFU_BOX_WARPS;
  const int st = tid - 32 * EFU_BOX_WARPS;   // index among the stream threads
  const int q3 = (C + 2) / 3;
  if (tid == 32 * EFU_BOX_WARPS) {
#pragma unroll
    for (int s = 0; s < EFU_STAGES; ++s) {
      const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (t < n_tiles) issue(t, s);
    }
  }
  uint32_t phase_bits = 0;   // parity of each stage's barrier
  int j = 0;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++j) {
    const int s = j & (EFU_STAGES - 1), pb = j & 1;
    int l, nrec; long long rec0;
    locate(t, l, rec0, nrec);
    if (WITH_LOSS && l != cur_level) {
      if (cur_level >= 0) flush(cur_level);
      cur_level = l;
    }
    const int api = p.lv.anc_per_img[l];
    const long long total = (long long)p.B * api;
    const bool bulk = bulk_ok(l, rec0, nrec);
    int img = 0;
    uint32_t aidx = 0;
    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (box_warp) {
      // ---- box half: lane <-> anchor ----
      if (tid < nrec) {
        const long long rec = rec0 + tid;
        const float4 r4 = __ldcs(p.rel[l] + rec);
        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned char mk = 0;
        if (WITH_LOSS) { t4 = __ldcs(p.box_true[l] + rec); mk = p.mask[l][rec]; }
        int rin;
        if (total < 0x7fffffffLL) { img = (int)((uint32_t)rec / (uint32_t)api); rin = (int)((uint32_t)rec - (uint32_t)img * (uint32_t)api); }
        else { img = (int)(rec / api); rin = (int)(rec - (long long)img * api); }
        int y, x, an_i;
        ef_split(p.lv, l, rin, y, x, an_i);
        const AnchorBox an = ef_anchor(p.lv, l, y, x, an_i);
        // _boxes_decoder, anc:245-274 (the arithmetic of effdet_decode_kernel)
        const float yca = DM_DIV(DM_ADD(an.y2, an.y1), 2.0f), xca = DM_DIV(DM_ADD(an.x2, an.x1), 2.0f);
        const float ha = DM_SUB(an.y2, an.y1), wa = DM_SUB(an.x2, an.x1);
        const float w = DM_MUL(dm_expf(r4.w), wa), h = DM_MUL(dm_expf(r4.z), ha);
        const float yc = DM_ADD(DM_MUL(r4.x, ha), yca), xc = DM_ADD(DM_MUL(r4.y, wa), xca);
        const float hh = DM_DIV(h, 2.0f), hw = DM_DIV(w, 2.0f);
        d4 = make_float4(DM_SUB(yc, hh), DM_SUB(xc, hw), DM_ADD(yc, hh), DM_ADD(xc, hw));
        if (p.dec[l]) __stcs(p.dec[l] + rec, d4);
        if (WITH_LOSS) {
          hub += el_huber(t4.x, r4.x, p.delta) + el_huber(t4.y, r4.y, p.delta) + el_huber(t4.z, r4.z, p.delta) + el_huber(t4.w, r4.w, p.delta);
          pos += mk ? 1u : 0u;
        }
        aidx = (uint32_t)(p.lv.anchor_base[l] + rin);
      }
    } else {
      // ---- class stream ----
      float* xs = reinterpret_cast<float*>(efu_smem + (size_t)s * stage_bytes);
      const float* ys = reinterpret_cast<const float*>(efu_smem + (size_t)s * stage_bytes + tile_bytes);
      const int n_el = nrec * C;
      float f0 = 0.f, f1 = 0.f;
      if (bulk) {
        bc_mbar_wait(&s_bar[s], (phase_bits >> s) & 1u);
        if (WITH_LOSS) {
          const int n_vec = n_el >> 2;
#pragma unroll 2
          for (int i = st; i < n_vec; i += EFU_STREAM_THREADS) {
            const float4 x = reinterpret_cast<const float4*>(xs)[i];
            const float4 y = reinterpret_cast<const float4*>(ys)[i];
            f0 += el_focal<G15>(y.x, x.x, p.alpha, p.gamma, p.label_smoothing) + el_focal<G15>(y.y, x.y, p.alpha, p.gamma, p.label_smoothing);
            f1 += el_focal<G15>(y.z, x.z, p.alpha, p.gamma, p.label_smoothing) + el_focal<G15>(y.w, x.w, p.alpha, p.gamma, p.label_smoothing);
          }
        }
      } else {
        // ragged tail of a level or an unaligned tensor: plain loads; the logits still go to the stage for the argmax
        const float* xg = p.cls[l] + rec0 * C;
        const float* yg = WITH_LOSS ? p.cls_true[l] + rec0 * C : nullptr;
        for (int e = st; e < n_el; e += EFU_STREAM_THREADS) {
          const float xv = __ldg(xg + e);
          xs[e] = xv;
          if (WITH_LOSS) f0 += el_focal<G15>(__ldg(yg + e), xv, p.alpha, p.gamma, p.label_smoothing);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EFU_STREAM_THREADS) : "memory");   // stream warps only: the tile is complete
      }
      if (WITH_LOSS) focal += (double)f0 + (double)f1;
      // per-anchor argmax, three lanes per anchor: tf.argmax (first maximal index) / reduce_max over a third of the C
      // logits each; the box warps fold the three partials in ascending index order
      const int a = st / 3, part = st - a * 3;
      float m = -INFINITY;
      int mi = -1;
      if (a < nrec) {
        const float* r = xs + a * C;
        const int c0 = part * q3, c1 = min(C, c0 + q3);
        int c = c0;
        if (part == 0) { m = r[0]; mi = 0; c = 1; }   // the scan starts from element 0 exactly as a serial one (NaN there sticks)
        for (; c + 4 <= c1; c += 4) {                  // loads first, then the dependent compare / select chain
          const float v0 = r[c], v1 = r[c + 1], v2 = r[c + 2], v3 = r[c + 3];
          if (v0 > m) { m = v0; mi = c; }
          if (v1 > m) { m = v1; mi = c + 1; }
          if (v2 > m) { m = v2; mi = c + 2; }
          if (v3 > m) { m = v3; mi = c + 3; }
        }
        for (; c < c1; ++c) { const float v = r[c]; if (v > m) { m = v; mi = c; } }
      }
      s_m[pb][a][part] = m;
      s_mi[pb][a][part] = mi;
    }
    __syncthreads();   // the tile's single CTA barrier: the stage is consumed, the argmax partials are published
    if (bulk) phase_bits ^= (1u << s);
    if (tid == 32 * EFU_BOX_WARPS) {
      const long long tn = t + (long long)EFU_STAGES * gridDim.x;
      if (tn < n_tiles) issue(tn, s);
    }
    if (box_warp) {
      // ---- fold the partials, candidate append (anc:179-189), one atomic per (warp, image) ----
      float m = s_m[pb][tid][0];
      int mi = s_mi[pb][tid][0];
#pragma unroll
      for (int k = 1; k < 3; ++k) { const float vk = s_m[pb][tid][k]; if (vk > m) { m = vk; mi = s_mi[pb][tid][k]; } }
      const bool pass = (tid < nrec) && (mi != 0);   // classes_mask = classes_id != 0 (anc:179)
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
          p.cand_box[slot] = d4; p.cand_score[slot] = m; p.cand_cls[slot] = mi; p.cand_aidx[slot] = aidx;
          if (p.bitmap) atomicOr(&p.bitmap[(size_t)limg * p.bitmap_words + (aidx >> 5)], 1u << (aidx & 31u));
        }
        todo &= ~grp;
      }
    }
    // the partial buffers alternate by tile parity: the stream warps write buffer pb again only two tiles later, behind
    // the next tile's barrier, which the box warps reach after this append
  }
  if (WITH_LOSS) {
    if (cur_level >= 0) flush(cur_level);
    __syncthreads();
    if (tid < p.lv.num_levels * 3)
      p.partials[((size_t)blockIdx.x * p.lv.num_levels + tid / 3) * 3 + tid % 3] = s_part[tid / 3][tid % 3];
  }
}

// sums layout of effdet_loss.cu: [0..L) focal_l, [L..2L) huber_l, [2L] positives
__global__ void __launch_bounds__(256) effdet_stream_reduce_kernel(const double* __restrict__ partials, int n_cta, int L, double* __restrict__ sums) {
  __shared__ double s_red[8][3];
  const int l = blockIdx.x;
  double a[3] = {0.0, 0.0, 0.0};
  for (int c = threadIdx.x; c < n_cta; c += 256) {
    const double* q = partials + ((size_t)c * L + l) * 3;
    a[0] += q[0]; a[1] += q[1]; a[2] += q[2];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) { const double v = warp_sum_d(a[k]); if (lane == 0) s_red[warp][k] = v; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sacc[3] = {0.0, 0.0, 0.0};
    for (int w = 0; w < 8; ++w) for (int k = 0; k < 3; ++k) sacc[k] += s_red[w][k];
    sums[l] = sacc[0];
    sums[L + l] = sacc[1];
    atomicAdd(&sums[2 * L], sacc[2]);  // integer-valued: exact and order-independent
  }
}

#define EFU_GRID_PER_SM 16

static size_t efu_partials_bytes(int num_levels) {
  return b200_align_up(sizeof(double) * 3 * (size_t)num_levels * (size_t)b200_sm_count() * EFU_GRID_PER_SM, 256);
}

extern "C" size_t b200_effdet_eval_workspace_bytes(int num_levels, const int32_t* hw, int A, int num_images, int max_out) {
  return b200_effdet_postprocess_workspace_bytes(num_levels, hw, A, num_images, max_out) + efu_partials_bytes(num_levels);
}

static int efu_impl(bool with_loss, int num_levels, const int32_t* hw, int A, const float* table_dev, int C, int B,
                    const float* const true_boxes[], const float* const true_classes[], const unsigned char* const true_masks[],
                    const float* const pred_boxes[], const float* const pred_classes[], float alpha, float gamma, float delta,
                    float label_smoothing, double* sums_out, float* const out_decoded[], int max_out, float iou_thr,
                    float score_thr, int metric, float* out_boxes, long long* out_class_id, float* out_score,
                    int32_t* out_sel_idx, int32_t* out_sel_anchor, int32_t* out_count, void* workspace, size_t workspace_bytes,
                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const char* who = with_loss ? "b200_effdet_eval_step" : "b200_effdet_decode_postprocess";
  EfFusedParams p;
  B200_REQUIRE(hw && table_dev && pred_boxes && pred_classes, B200_ERR_BAD_ARG, "%s: null argument", who);
  B200_REQUIRE(!with_loss || (true_boxes && true_classes && true_masks && sums_out), B200_ERR_BAD_ARG, "%s: null target argument", who);
  const int n_img = ef_fill_levels(p.lv, num_levels, hw, A, table_dev);
  B200_REQUIRE(n_img >= 0, B200_ERR_BAD_ARG, "%s: bad level spec", who);
  B200_REQUIRE(C >= 1 && C <= 400, B200_ERR_UNSUPPORTED, "%s: classes_num %d outside [1,400]", who, C);
  B200_REQUIRE(B >= 0, B200_ERR_BAD_ARG, "%s: negative batch", who);
  B200_REQUIRE(metric >= B200_METRIC_EFF_IOU && metric <= B200_METRIC_EFF_CIOU, B200_ERR_BAD_ARG, "%s: iou_type must be iou/giou/diou/ciou", who);
  B200_REQUIRE(max_out >= 1 && max_out <= NMS_MAX_OUT_LIMIT, B200_ERR_UNSUPPORTED, "%s: max_out %d outside [1,%d]", who, max_out, NMS_MAX_OUT_LIMIT);
  if (B == 0) return B200_OK;
  B200_REQUIRE(out_boxes && out_class_id && out_score && out_count, B200_ERR_BAD_ARG, "%s: null output", who);
  EfWs ws = ef_ws_layout(B, n_img, max_out);
  const size_t part_bytes = efu_partials_bytes(num_levels);
  B200_REQUIRE(workspace && workspace_bytes >= ws.total + part_bytes, B200_ERR_WORKSPACE, "%s: workspace %zu < required %zu", who, workspace_bytes, ws.total + part_bytes);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, B200_ERR_BAD_ARG, "%s: workspace not 256-byte aligned", who);
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  p.B = B; p.C = C; p.n_img = n_img;
  long long tb = 0;
  for (int l = 0; l < EF_MAX_LEVELS; ++l) {
    p.tile_base[l] = tb;
    p.cls[l] = nullptr; p.cls_true[l] = nullptr; p.rel[l] = nullptr; p.box_true[l] = nullptr; p.mask[l] = nullptr; p.dec[l] = nullptr;
    if (l >= num_levels) continue;
    B200_REQUIRE(pred_boxes[l] && pred_classes[l], B200_ERR_BAD_ARG, "%s: null level %d", who, l);
    B200_REQUIRE(!with_loss || (true_boxes[l] && true_classes[l] && true_masks[l]), B200_ERR_BAD_ARG, "%s: null target level %d", who, l);
    uintptr_t al = reinterpret_cast<uintptr_t>(pred_boxes[l]) | (out_decoded && out_decoded[l] ? reinterpret_cast<uintptr_t>(out_decoded[l]) : 0);
    if (with_loss) al |= reinterpret_cast<uintptr_t>(true_boxes[l]);
    B200_REQUIRE((al & 15) == 0, B200_ERR_BAD_ARG, "%s: box tensors of level %d must be 16-byte aligned", who, l);
    p.cls[l] = pred_classes[l];
    p.rel[l] = reinterpret_cast<const float4*>(pred_boxes[l]);
    p.dec[l] = out_decoded ? reinterpret_cast<float4*>(out_decoded[l]) : nullptr;
    if (with_loss) {
      p.cls_true[l] = true_classes[l];
      p.box_true[l] = reinterpret_cast<const float4*>(true_boxes[l]);
      p.mask[l] = true_masks[l];
    }
    tb += ((long long)B * p.lv.anc_per_img[l] + EFU_TILE - 1) / EFU_TILE;
  }
  for (int l = num_levels; l <= EF_MAX_LEVELS; ++l) p.tile_base[l] = tb;
  p.cand_box = reinterpret_cast<float4*>(wsb + ws.box);
  p.cand_score = reinterpret_cast<float*>(wsb + ws.score);
  p.cand_cls = reinterpret_cast<int32_t*>(wsb + ws.cls);
  p.cand_aidx = reinterpret_cast<uint32_t*>(wsb + ws.aidx);
  p.counts = reinterpret_cast<int32_t*>(wsb + ws.counts);
  p.bitmap = out_sel_idx ? reinterpret_cast<uint32_t*>(wsb + ws.bitmap) : nullptr;
  p.bitmap_words = ws.bitmap_words;
  p.alpha = alpha; p.gamma = gamma; p.delta = delta; p.label_smoothing = label_smoothing;
  p.partials = reinterpret_cast<double*>(wsb + ws.total);
  B200_CUDA(cudaMemsetAsync(wsb, 0, out_sel_idx ? ws.box : ws.bitmap, stream));
  B200_REQUIRE(C <= 128, B200_ERR_UNSUPPORTED, "%s: the fused pass stages %d-anchor tiles in shared memory: classes_num %d > 128 (use the separate calls)", who, EFU_TILE, C);
  const size_t smem = (size_t)EFU_STAGES * EFU_TILE * C * sizeof(float) * (with_loss ? 2 : 1);
  long long want = tb;
  const long long cap = (long long)b200_sm_count() * EFU_GRID_PER_SM;
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
#define EFU_LAUNCH(WL, G)                                                                                                  \
  do {                                                                                                                     \
    B200_CUDA(cudaFuncSetAttribute(effdet_stream_kernel<WL, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    effdet_stream_kernel<WL, G><<<grid, EFU_THREADS, smem, stream>>>(p);                                                  \
  } while (0)
  B200_REQUIRE(with_loss, B200_ERR_BAD_ARG, "%s: the CTA-tiled stream needs targets (the target-free pass is the bulk-copy filter)", who);
  if (gamma == 1.5f) EFU_LAUNCH(true, true);
  else EFU_LAUNCH(true, false);
#undef EFU_LAUNCH
  B200_LAUNCH_CHECK();
  if (with_loss) {
    B200_CUDA(cudaMemsetAsync(sums_out, 0, sizeof(double) * (2 * num_levels + 1), stream));
    effdet_stream_reduce_kernel<<<num_levels, 256, 0, stream>>>(p.partials, grid, num_levels, sums_out);
    B200_LAUNCH_CHECK();
  }
  // the NMS tail reads the candidate store through the filter-parameter view
  EfFilterParams fp;
  fp.lv = p.lv; fp.B0 = 0; fp.NB = B; fp.C = C; fp.B = B; fp.n_img = n_img;
  fp.cand_box = p.cand_box; fp.cand_score = p.cand_score; fp.cand_cls = p.cand_cls; fp.cand_aidx = p.cand_aidx;
  fp.counts = p.counts; fp.bitmap = p.bitmap; fp.bitmap_words = p.bitmap_words;
  return ef_launch_nms(fp, ws, wsb, B, n_img, max_out, iou_thr, score_thr, metric, out_boxes, out_class_id, out_score, out_sel_idx,
                       out_sel_anchor, out_count, stream);
}

// test_step in one call: sums_out [2L+1] fp64 (un-normalised focal_l, huber_l, positives: feed b200_focal_box_finalize or
// its _dp form), out_decoded[l] = convert_outputs_boxes, then convert_outputs_one for every image of the batch.
extern "C" int b200_effdet_eval_step(int num_levels, const int32_t* hw, int A, const float* table_dev, int C, int B,
                                     const float* const true_boxes[], const float* const true_classes[],
                                     const unsigned char* const true_masks[], const float* const pred_boxes[],
                                     const float* const pred_classes[], float alpha, float gamma, float delta,
                                     float label_smoothing, double* sums_out, float* const out_decoded[], int max_out,
                                     float iou_thr, float score_thr, int metric, float* out_boxes, long long* out_class_id,
                                     float* out_score, int32_t* out_sel_idx, int32_t* out_sel_anchor, int32_t* out_count,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  return efu_impl(true, num_levels, hw, A, table_dev, C, B, true_boxes, true_classes, true_masks, pred_boxes, pred_classes, alpha,
                  gamma, delta, label_smoothing, sums_out, out_decoded, max_out, iou_thr, score_thr, metric, out_boxes, out_class_id,
                  out_score, out_sel_idx, out_sel_anchor, out_count, workspace, workspace_bytes, stream);
}

// convert_outputs_boxes + convert_outputs_one of the whole batch in one pass over the heads (no targets, no loss).
extern "C" int b200_effdet_decode_postprocess(int num_levels, const int32_t* hw, int A, const float* table_dev, int C, int B,
                                              const float* const pred_boxes[], const float* const pred_classes[],
                                              float* const out_decoded[], int max_out, float iou_thr, float score_thr, int metric,
                                              float* out_boxes, long long* out_class_id, float* out_score, int32_t* out_sel_idx,
                                              int32_t* out_sel_anchor, int32_t* out_count, void* workspace, size_t workspace_bytes,
                                              void* stream) {
  // without targets the per-warp bulk-copy filter (lane <-> anchor, no CTA barriers) is the faster stream: the decode rides
  // on it (one coalesced 16-byte load and store per lane); the CTA-tiled kernel above pays off only with the focal terms
  B200_REQUIRE(pred_boxes && table_dev, B200_ERR_BAD_ARG, "b200_effdet_decode_postprocess: null argument");
  return ef_post_impl(num_levels, hw, A, table_dev, C, B, 0, B, nullptr, pred_boxes, out_decoded, pred_classes, max_out, iou_thr,
                      score_thr, metric, out_boxes, out_class_id, out_score, out_sel_idx, out_sel_anchor, out_count, workspace,
                      workspace_bytes, stream);
}
