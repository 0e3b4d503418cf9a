// nms.cuh — exact greedy NMS for one segment (image) per CTA.  Included by nms.cu (generic C-ABI entry),
// yolo_decode.cu and effdet.cu (fused post-processing).
//
// Replaces the tf.while_loop NMS drivers of the reference:
//   utils/tf_iou_utils.py:67-108   GetIOUNMS            (class-agnostic, survivor iff metric <  thr; NaN drops)
//   utils/tf_iou_utils.py:110-157  GetIOUNMSByClasses   (suppressed iff metric >= thr and same class; NaN stays)
//   efficientnet/utils/nms.py:5-61 get_nms              (GetIOUNMS + stop at first top score < score_threshold)
// Semantics reproduced exactly: candidates ordered by (score descending, index ascending) as tf.argsort
// (DESCENDING) does; candidate j is emitted iff no earlier *emitted* i suppresses it; at most max_out emitted.
//
// Algorithm (B200: one 1024-thread CTA per image, everything after the score read lives in shared memory):
//   1. radix-select, 8 bits per pass on the 64-bit key (~ordered(score) << 32 | order_id), the next chunk of
//      <= 1024 candidates in sorted order (per-warp shared-memory histograms; usually 2 passes);
//   2. bitonic sort of the chunk (64-bit keys, position payload);
//   3. suppression in tiles of 64 sorted candidates: 64x64 bitmask built with warp ballots, a one-warp
//      sweep over the tile (skipped when the tile has no internal conflicts), then every later candidate
//      of the chunk tests itself against the boxes emitted by this tile.  Work is O(emitted x chunk), not
//      O(chunk^2), and stops as soon as max_out boxes are out;
//   4. if fewer than max_out were emitted and candidates remain, select the next chunk; its candidates are
//      first tested against all boxes emitted so far (kept list in shared memory).
#pragma once
#include "boxmath.cuh"
#include "common.cuh"

#define NMS_THREADS 1024
#define NMS_CHUNK 1024
#define NMS_MIN_FILL 192
#define NMS_MAX_OUT_LIMIT 4096

struct NmsSegment {
  const float* boxes;        // [n,4] in the metric family's corner order
  const float* scores;       // [n]
  const int32_t* classes;    // [n] or nullptr (required for BY_CLASS)
  const uint32_t* order_id;  // [n] unique tie-break ids (ascending = earlier) or nullptr -> position
  int n;
};

struct NmsConfig {
  int metric;
  int mode;
  float iou_thr;
  float score_thr;
  int use_score_thr;
  int max_out;
};

__host__ __device__ inline size_t nms_smem_bytes(int max_out) {
  // sK 8K + sPos 4K + cand 28K + hist 32K + kept 28B*max_out + small
  return (size_t)NMS_CHUNK * 8 + NMS_CHUNK * 4 + NMS_CHUNK * 28 + 32 * 256 * 4 + (size_t)max_out * 28 + 1024;
}

__device__ __forceinline__ uint32_t nms_dkey(float s) {
  uint32_t u = __float_as_uint(s);
  if ((u << 1) == 0u) u = 0u;  // -0 == +0 for tf.argsort
  uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~k;  // ascending dkey == descending score
}

__device__ __forceinline__ bool nms_suppresses(const BoxT& kept, int kept_cls, const BoxT& cand, int cand_cls,
                                                const NmsConfig& cfg) {
  if (cfg.mode == B200_NMS_BY_CLASS) {
    if (kept_cls != cand_cls) return false;
    if (bm_surely_below(kept, cand, cfg.metric, cfg.iou_thr)) return false;
    float m = bm_metric(kept, cand, cfg.metric);
    return m >= cfg.iou_thr;
  }
  if (bm_surely_below(kept, cand, cfg.metric, cfg.iou_thr)) return false;
  float m = bm_metric(kept, cand, cfg.metric);
  return !(m < cfg.iou_thr);
}

// Runs NMS for one segment with the whole CTA (blockDim.x == NMS_THREADS).  out_pos receives the local
// positions (0..n-1) of emitted candidates in emit order; returns the number emitted (uniform across threads).
static __device__ int nms_run_segment(const NmsSegment& seg, const NmsConfig& cfg, int32_t* __restrict__ out_pos,
                               unsigned char* smem_raw) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  unsigned long long* sK = reinterpret_cast<unsigned long long*>(smem_raw);
  uint32_t* sPos = reinterpret_cast<uint32_t*>(sK + NMS_CHUNK);
  float* cC0 = reinterpret_cast<float*>(sPos + NMS_CHUNK);
  float* cC1 = cC0 + NMS_CHUNK;
  float* cC2 = cC1 + NMS_CHUNK;
  float* cC3 = cC2 + NMS_CHUNK;
  float* cAr = cC3 + NMS_CHUNK;
  float* cAt = cAr + NMS_CHUNK;
  int* cCl = reinterpret_cast<int*>(cAt + NMS_CHUNK);
  uint32_t* sHist = reinterpret_cast<uint32_t*>(cCl + NMS_CHUNK);  // [32][256]
  float* kC0 = reinterpret_cast<float*>(sHist + 32 * 256);
  float* kC1 = kC0 + cfg.max_out;
  float* kC2 = kC1 + cfg.max_out;
  float* kC3 = kC2 + cfg.max_out;
  float* kAr = kC3 + cfg.max_out;
  float* kAt = kAr + cfg.max_out;
  int* kCl = reinterpret_cast<int*>(kAt + cfg.max_out);
  // small scalars after the kept arrays (8-byte aligned: max_out*28 is a multiple of 4; pad)
  uintptr_t misc_addr = (reinterpret_cast<uintptr_t>(kCl + cfg.max_out) + 15) & ~uintptr_t(15);
  unsigned long long* sMask = reinterpret_cast<unsigned long long*>(misc_addr);  // [64]
  uint32_t* sAlive = reinterpret_cast<uint32_t*>(sMask + 64);                    // [32]
  uint32_t* sWarpTot = sAlive + 32;                                             // [32]
  unsigned long long* sKeptMask = reinterpret_cast<unsigned long long*>(sWarpTot + 32);
  int* sScalar = reinterpret_cast<int*>(sKeptMask + 1);  // [0]=gather count [1]=accepted add [2]=nk

  const int n = seg.n;
  int n_kept = 0;
  unsigned long long klo = 0ull;  // inclusive lower bound of not-yet-consumed keys
  bool exhausted = (n <= 0);

  while (!exhausted && n_kept < cfg.max_out) {
    // ---------------- 1. choose khi by radix descent ----------------
    unsigned long long base = 0ull;  // start of the current bucket
    int count_below = 0;             // eligible keys in [klo, base)
    bool last = false;               // true -> everything remaining fits in this chunk
    unsigned long long khi = 0ull;
    for (int level = 0; level < 8; ++level) {
      const int shift = 56 - 8 * level;
      for (int i = tid; i < 32 * 256; i += NMS_THREADS) sHist[i] = 0u;
      __syncthreads();
      for (int i = tid; i < n; i += NMS_THREADS) {
        float s = seg.scores[i];
        if (cfg.use_score_thr && (s < cfg.score_thr)) continue;
        uint32_t oid = seg.order_id ? seg.order_id[i] : (uint32_t)i;
        unsigned long long K = ((unsigned long long)nms_dkey(s) << 32) | oid;
        if (K < klo) continue;
        if (level > 0 && ((K >> (shift + 8)) != (base >> (shift + 8)))) continue;  // outside current bucket
        atomicAdd(&sHist[warp * 256 + (int)((K >> shift) & 255ull)], 1u);
      }
      __syncthreads();
      // reduce the 32 per-warp histograms and scan the 256 bins
      int h = 0;
      if (tid < 256) {
#pragma unroll 8
        for (int w = 0; w < 32; ++w) h += (int)sHist[w * 256 + tid];
      }
      int incl = warp_scan_incl(h);
      if (tid < 256 && lane == 31) sWarpTot[warp] = (uint32_t)incl;
      __syncthreads();
      if (tid < 256) {
        int off = 0;
        for (int w = 0; w < warp; ++w) off += (int)sWarpTot[w];
        incl += off;
      }
      int fits = (tid < 256) && (count_below + incl <= NMS_CHUNK);
      int m = __syncthreads_count(fits);  // largest m with count_below + sum_{d<m} h[d] <= CHUNK
      if (tid == 0) sScalar[1] = 0;
      __syncthreads();
      if (tid < 256 && tid == m - 1) sScalar[1] = incl;
      __syncthreads();
      count_below += sScalar[1];
      if (m == 256) {
        if (level == 0) { last = true; break; }
        // whole bucket fits after all (cannot happen: its parent bin overflowed); treat as boundary
        khi = base + (1ull << (shift + 8));
        break;
      }
      // bins < m of this level are accepted; bin m overflows and becomes the next bucket
      base += ((unsigned long long)m << shift);
      khi = base;
      if (count_below >= NMS_MIN_FILL || level == 7) {
        if (count_below == 0) khi = base + 1ull;  // duplicate keys (invalid order ids): force progress
        break;
      }
      __syncthreads();
    }
    __syncthreads();

    // ---------------- 2. gather + sort the chunk ----------------
    if (tid == 0) sScalar[0] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NMS_THREADS) {
      float s = seg.scores[i];
      if (cfg.use_score_thr && (s < cfg.score_thr)) continue;
      uint32_t oid = seg.order_id ? seg.order_id[i] : (uint32_t)i;
      unsigned long long K = ((unsigned long long)nms_dkey(s) << 32) | oid;
      if (K < klo) continue;
      if (!last && K >= khi) continue;
      int slot = atomicAdd(&sScalar[0], 1);
      if (slot < NMS_CHUNK) { sK[slot] = K; sPos[slot] = (uint32_t)i; }
    }
    __syncthreads();
    int n_chunk = min(sScalar[0], NMS_CHUNK);
    if (n_chunk == 0) { exhausted = true; break; }
    int npad = 64;
    while (npad < n_chunk) npad <<= 1;
    if (tid >= n_chunk && tid < npad) { sK[tid] = ~0ull; sPos[tid] = 0xffffffffu; }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        int ixj = tid ^ j;
        if (tid < npad && ixj > tid) {
          unsigned long long a = sK[tid], b = sK[ixj];
          bool asc = ((tid & k) == 0);
          if ((a > b) == asc) {
            sK[tid] = b; sK[ixj] = a;
            uint32_t pa = sPos[tid]; sPos[tid] = sPos[ixj]; sPos[ixj] = pa;
          }
        }
        __syncthreads();
      }
    }

    // ---------------- 3. load candidates (sorted order: thread t owns sorted candidate t) ----------------
    BoxT mine; int my_cls = 0; bool alive = false;
    mine.c0 = mine.c1 = mine.c2 = mine.c3 = mine.area = mine.at = 0.0f;
    if (tid < n_chunk) {
      uint32_t p = sPos[tid];
      float4 b = __ldg(reinterpret_cast<const float4*>(seg.boxes) + p);
      mine = bm_prep(b.x, b.y, b.z, b.w, cfg.metric);
      my_cls = seg.classes ? seg.classes[p] : 0;
      alive = true;
      cC0[tid] = mine.c0; cC1[tid] = mine.c1; cC2[tid] = mine.c2; cC3[tid] = mine.c3;
      cAr[tid] = mine.area; cAt[tid] = mine.at; cCl[tid] = my_cls;
    }
    // phase A: against everything emitted by earlier chunks
    if (alive) {
      for (int k = 0; k < n_kept; ++k) {
        BoxT kb; kb.c0 = kC0[k]; kb.c1 = kC1[k]; kb.c2 = kC2[k]; kb.c3 = kC3[k]; kb.area = kAr[k]; kb.at = kAt[k];
        if (nms_suppresses(kb, kCl[k], mine, my_cls, cfg)) { alive = false; break; }
      }
    }
    {
      uint32_t bal = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) sAlive[warp] = bal;
    }
    __syncthreads();

    // ---------------- 4. tiles of 64 ----------------
    const int n_tiles = (n_chunk + 63) >> 6;
    for (int T = 0; T < n_tiles && n_kept < cfg.max_out; ++T) {
      const int t0 = T << 6;
      const unsigned long long tile_alive =
          (unsigned long long)sAlive[2 * T] | ((unsigned long long)sAlive[2 * T + 1] << 32);
      // 4a. intra-tile mask via ballots: warp w owns rows 2w, 2w+1
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = 2 * warp + rr;  // row within tile
        const int gi = t0 + r;
        const bool row_on = (gi < n_chunk) && ((tile_alive >> r) & 1ull);
        BoxT rb; int rcls = 0;
        if (row_on) { rb.c0 = cC0[gi]; rb.c1 = cC1[gi]; rb.c2 = cC2[gi]; rb.c3 = cC3[gi]; rb.area = cAr[gi]; rb.at = cAt[gi]; rcls = cCl[gi]; }
        uint32_t bits[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c = half * 32 + lane;
          const int gj = t0 + c;
          bool pred = false;
          if (row_on && c > r && gj < n_chunk && ((tile_alive >> c) & 1ull)) {
            BoxT cb; cb.c0 = cC0[gj]; cb.c1 = cC1[gj]; cb.c2 = cC2[gj]; cb.c3 = cC3[gj]; cb.area = cAr[gj]; cb.at = cAt[gj];
            pred = nms_suppresses(rb, rcls, cb, cCl[gj], cfg);
          }
          bits[half] = __ballot_sync(0xffffffffu, pred);
        }
        if (lane == 0) sMask[r] = (unsigned long long)bits[0] | ((unsigned long long)bits[1] << 32);
      }
      __syncthreads();
      // 4b. one-warp sweep
      if (warp == 0) {
        unsigned long long m0 = ((tile_alive >> lane) & 1ull) ? (sMask[lane] & tile_alive) : 0ull;
        unsigned long long m1 = ((tile_alive >> (lane + 32)) & 1ull) ? (sMask[lane + 32] & tile_alive) : 0ull;
        bool any = __any_sync(0xffffffffu, (m0 | m1) != 0ull);
        unsigned long long keptmask = tile_alive;
        const int room = cfg.max_out - n_kept;
        if (lane == 0) {
          if (any) {
            unsigned long long rem = tile_alive;
            keptmask = 0ull;
            int nk = 0;
            while (rem && nk < room) {
              int i = __ffsll((long long)rem) - 1;
              keptmask |= (1ull << i);
              ++nk;
              rem &= ~(1ull << i);
              rem &= ~sMask[i];
            }
          } else {
            int nk = __popcll(keptmask);
            while (nk > room) {  // keep only the first `room` alive candidates
              int hi = 63 - __clzll((long long)keptmask);
              keptmask &= ~(1ull << hi);
              --nk;
            }
          }
          *sKeptMask = keptmask;
          sScalar[2] = __popcll(keptmask);
        }
      }
      __syncthreads();
      const unsigned long long keptmask = *sKeptMask;
      const int nk = sScalar[2];
      // 4c. emit + cross-tile suppression
      if (tid >= t0 && tid < t0 + 64) {
        const int r = tid - t0;
        if ((keptmask >> r) & 1ull) {
          int slot = n_kept + __popcll(keptmask & ((1ull << r) - 1ull));
          kC0[slot] = mine.c0; kC1[slot] = mine.c1; kC2[slot] = mine.c2; kC3[slot] = mine.c3;
          kAr[slot] = mine.area; kAt[slot] = mine.at; kCl[slot] = my_cls;
          out_pos[slot] = (int32_t)sPos[tid];
        }
        alive = false;  // tile resolved
      } else if (tid >= t0 + 64 && alive) {
        unsigned long long km = keptmask;
        while (km) {
          int i = __ffsll((long long)km) - 1;
          km &= km - 1ull;
          const int gi = t0 + i;
          BoxT kb; kb.c0 = cC0[gi]; kb.c1 = cC1[gi]; kb.c2 = cC2[gi]; kb.c3 = cC3[gi]; kb.area = cAr[gi]; kb.at = cAt[gi];
          if (nms_suppresses(kb, cCl[gi], mine, my_cls, cfg)) { alive = false; break; }
        }
      }
      n_kept += nk;
      {
        uint32_t bal = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) sAlive[warp] = bal;
      }
      __syncthreads();
    }
    if (last) exhausted = true;
    klo = khi;
    __syncthreads();
  }
  return n_kept;
}
