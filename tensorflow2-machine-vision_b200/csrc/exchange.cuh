// exchange.cuh — the one collective of the path: the sum of a handful of per-GPU partial loss terms over the ranks of a
// data-parallel job (the reference's only collective site is the Mirrored-strategy reduce_sum of
// facenet/facenet_model.py:297,318-322).
//
// B200-first form: no separate collective launch.  Every rank owns a small "mailbox" in its HBM that its peers map
// through CUDA IPC (NVLink 5 / NVSwitch peer stores).  The last CTA of the loss-finalize kernel
//   1. lane r stores the n <= 32 partial terms into slot [epoch % 4][rank] of rank r's mailbox (plain peer stores) and
//   2. release-stores the epoch number into that slot's flag (the release orders the lane's own stores: no wider fence),
//   3. lane r acquire-spins on rank r's flag in its OWN mailbox (local HBM, no link traffic while waiting),
//   4. the warp adds the `world` payloads in rank order — the same order on every rank, so every rank gets the same bits.
// Four slot sets are used round robin.  Fused form: a rank can be at most one exchange ahead of its slowest peer (it
// needs that peer's flag of epoch e to leave epoch e), so a set is never overwritten while somebody still reads it.
// Split form (publish now, collect later on another stream): rule "publish(f) is ordered after this rank's own
// collect(f - 2)".  Then rank A's publish(f) implies A collected f-2, hence peer B published f-2, hence (B's rule) B
// collected f-4 — exactly the epoch whose slot set publish(f) overwrites in B's mailbox.
// A wall-clock bound (%globaltimer) turns a missing peer into an error flag instead of a hung GPU.
#pragma once
#include <stdint.h>

#define B200_XCHG_MAX_WORLD 8
#define B200_XCHG_MAX_VALUES 32
#define B200_XCHG_SLOT_BYTES 512   // 32 x 8-byte payload + flag, padded
#define B200_XCHG_HEADER_BYTES 256
#define B200_XCHG_SLOT_SETS 4      // slot sets, used round robin by epoch
#define B200_XCHG_MAILBOX_BYTES (B200_XCHG_HEADER_BYTES + B200_XCHG_SLOT_SETS * B200_XCHG_MAX_WORLD * B200_XCHG_SLOT_BYTES)
#ifndef B200_XCHG_TIMEOUT_NS
#define B200_XCHG_TIMEOUT_NS 20000000000ull  // 20 s
#endif

// what a kernel needs to take part: passed by value inside the kernel's parameter struct (world == 1: no exchange)
struct B200Exchange {
  int rank, world;
  unsigned char* mailbox[B200_XCHG_MAX_WORLD];  // [r] = rank r's mailbox as mapped into THIS process (own one included)
};

#ifdef __CUDACC__
struct XchgHeader { unsigned long long epoch; unsigned int errors; unsigned int pad; unsigned long long pub_epoch; };  // epoch = exchanges collected

__device__ __forceinline__ unsigned char* xchg_slot(unsigned char* mailbox, unsigned par, int r) {
  return mailbox + B200_XCHG_HEADER_BYTES + ((size_t)par * B200_XCHG_MAX_WORLD + (size_t)r) * B200_XCHG_SLOT_BYTES;
}
__device__ __forceinline__ void xchg_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long xchg_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long xchg_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// The exchange in two halves, so that a caller may overlap the wait with other work (publish at the end of step i on
// the compute stream, collect on a second stream while step i+1 already runs).  Each half keeps its own epoch counter in
// the header; both advance by one per exchange.  The caller orders publish(f) after its own collect(f - 2) (stream /
// event order; see the slot-set argument at the top); the fused form below is trivially safe.
//
// publish: one value per lane (lanes >= n pass anything).  Lane r (< world) is the courier to rank r: it receives the n
// values by shuffles, stores them into slot [epoch % 4][rank] of rank r's mailbox and release-stores the flag behind them
// — the release of a thread orders that thread's own earlier stores, so no CTA- or system-wide fence is needed.
template <typename T>
__device__ __forceinline__ void xchg_publish_warp(const B200Exchange& x, T v, int n) {
  const int lane = threadIdx.x & 31;
  XchgHeader* hdr = reinterpret_cast<XchgHeader*>(x.mailbox[x.rank]);
  unsigned long long epoch = 0;
  if (lane == 0) epoch = *reinterpret_cast<volatile unsigned long long*>(&hdr->pub_epoch) + 1ull;
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const unsigned par = (unsigned)(epoch % B200_XCHG_SLOT_SETS);
  unsigned char* slot = lane < x.world ? xchg_slot(x.mailbox[lane], par, x.rank) : nullptr;   // own mailbox too
  for (int i = 0; i < n; ++i) {
    const T vi = __shfl_sync(0xffffffffu, v, i);
    if (slot) reinterpret_cast<volatile T*>(slot)[i] = vi;
  }
  if (slot) xchg_st_release_sys(reinterpret_cast<unsigned long long*>(slot + 8 * B200_XCHG_MAX_VALUES), epoch);
  if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&hdr->pub_epoch) = epoch;
}

// collect: lane r (< world) acquire-spins on rank r's flag of the next epoch in the LOCAL mailbox (no link traffic while
// waiting); the warp barrier then carries the happens-before edge to the lanes that add the payloads — in rank order, the
// same order on every rank.  Returns the sum on lanes < n.
template <typename T>
__device__ __forceinline__ T xchg_collect_warp(const B200Exchange& x, int n) {
  const int lane = threadIdx.x & 31;
  unsigned char* mine = x.mailbox[x.rank];
  XchgHeader* hdr = reinterpret_cast<XchgHeader*>(mine);
  unsigned long long epoch = 0;
  if (lane == 0) epoch = *reinterpret_cast<volatile unsigned long long*>(&hdr->epoch) + 1ull;
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const unsigned par = (unsigned)(epoch % B200_XCHG_SLOT_SETS);
  bool ok = true;
  if (lane < x.world) {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(xchg_slot(mine, par, lane) + 8 * B200_XCHG_MAX_VALUES);
    const unsigned long long t0 = xchg_globaltimer();
    while (xchg_ld_acquire_sys(flag) < epoch) {
      if (xchg_globaltimer() - t0 > B200_XCHG_TIMEOUT_NS) { ok = false; break; }
    }
  }
  ok = __all_sync(0xffffffffu, ok);   // also the barrier that orders the payload reads below behind every lane's acquire
  T acc = (T)0;
  if (lane < n) {
    for (int r = 0; r < x.world; ++r) {
      const volatile T* src = reinterpret_cast<const volatile T*>(xchg_slot(mine, par, r));
      acc += src[lane];
    }
  }
  if (lane == 0) {
    if (!ok) atomicAdd(&hdr->errors, 1u);
    *reinterpret_cast<volatile unsigned long long*>(&hdr->epoch) = epoch;
  }
  return acc;
}

// All-reduce(sum) of one value per lane (lanes >= n pass 0 and get 0 back) over the ranks of `x`, executed by ONE full
// warp of one CTA per rank.  T = float or double.  Returns the sum on every lane < n; world <= 1 returns v unchanged.
template <typename T>
__device__ __forceinline__ T xchg_allreduce_warp(const B200Exchange& x, T v, int n) {
  if (x.world <= 1) return v;
  xchg_publish_warp<T>(x, v, n);
  return xchg_collect_warp<T>(x, n);
}
#endif
