// exchange.cuh — the one collective of the path: the sum of a handful of per-GPU partial loss terms over the ranks of a
// data-parallel job (the reference's only collective site is the Mirrored-strategy reduce_sum of
// facenet/facenet_model.py:297,318-322).
//
// B200-first form: no separate collective launch.  Every rank owns a small "mailbox" in its HBM that its peers map
// through CUDA IPC (NVLink 5 / NVSwitch peer stores).  The last CTA of the loss-finalize kernel
//   1. stores its n <= 32 partial terms into slot [epoch & 1][rank] of EVERY rank's mailbox (plain peer stores),
//   2. fences (system scope) and release-stores the epoch number into the slot's flag,
//   3. acquire-spins on the `world` flags of its OWN mailbox (local HBM, no link traffic while waiting),
//   4. adds the `world` payloads in rank order — the same order on every rank, so every rank gets the same bits.
// Two slot sets alternate by epoch parity: a rank can be at most one exchange ahead of its slowest peer (it needs
// that peer's flag of epoch e to leave epoch e), so set (e & 1) is never overwritten while somebody still reads it.
// A wall-clock bound (%globaltimer) turns a missing peer into an error flag instead of a hung GPU.
#pragma once
#include <stdint.h>

#define B200_XCHG_MAX_WORLD 8
#define B200_XCHG_MAX_VALUES 32
#define B200_XCHG_SLOT_BYTES 512   // 32 x 8-byte payload + flag, padded
#define B200_XCHG_HEADER_BYTES 256
#define B200_XCHG_MAILBOX_BYTES (B200_XCHG_HEADER_BYTES + 2 * B200_XCHG_MAX_WORLD * B200_XCHG_SLOT_BYTES)
#ifndef B200_XCHG_TIMEOUT_NS
#define B200_XCHG_TIMEOUT_NS 20000000000ull  // 20 s
#endif

// what a kernel needs to take part: passed by value inside the kernel's parameter struct (world == 1: no exchange)
struct B200Exchange {
  int rank, world;
  unsigned char* mailbox[B200_XCHG_MAX_WORLD];  // [r] = rank r's mailbox as mapped into THIS process (own one included)
};

#ifdef __CUDACC__
struct XchgHeader { unsigned long long epoch; unsigned int errors; unsigned int pad; };

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

// All-reduce(sum) of one value per lane (lanes >= n pass 0 and get 0 back) over the ranks of `x`, executed by ONE full
// warp of one CTA per rank.  T = float or double.  Returns the sum on every lane < n; world <= 1 returns v unchanged.
template <typename T>
__device__ __forceinline__ T xchg_allreduce_warp(const B200Exchange& x, T v, int n) {
  if (x.world <= 1) return v;
  const int lane = threadIdx.x & 31;
  unsigned char* mine = x.mailbox[x.rank];
  XchgHeader* hdr = reinterpret_cast<XchgHeader*>(mine);
  unsigned long long epoch = 0;
  if (lane == 0) epoch = *reinterpret_cast<volatile unsigned long long*>(&hdr->epoch) + 1ull;
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const unsigned par = (unsigned)(epoch & 1ull);
  // 1. payload into every rank's mailbox (own one too: the sum below then reads all ranks the same way)
  if (lane < n) {
    for (int r = 0; r < x.world; ++r) {
      volatile T* dst = reinterpret_cast<volatile T*>(xchg_slot(x.mailbox[r], par, x.rank));
      dst[lane] = v;
    }
  }
  __threadfence_system();
  __syncwarp();
  // 2. publish
  if (lane < x.world)
    xchg_st_release_sys(reinterpret_cast<unsigned long long*>(xchg_slot(x.mailbox[lane], par, x.rank) + 8 * B200_XCHG_MAX_VALUES), epoch);
  // 3. wait for every rank's flag of this epoch in the local mailbox
  bool ok = true;
  if (lane < x.world) {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(xchg_slot(mine, par, lane) + 8 * B200_XCHG_MAX_VALUES);
    const unsigned long long t0 = xchg_globaltimer();
    while (xchg_ld_acquire_sys(flag) < epoch) {
      if (xchg_globaltimer() - t0 > B200_XCHG_TIMEOUT_NS) { ok = false; break; }
      __nanosleep(64);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  __threadfence_system();
  // 4. the same rank order everywhere
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
#endif
