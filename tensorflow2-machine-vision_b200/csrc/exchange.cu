// exchange.cu — host side of the loss all-reduce (SURVEY §8b `b200_allreduce_loss`, §8e).
//
// Two interchangeable transports behind the C ABI, both stream-ordered and CUDA-graph capturable:
//   * peer mailboxes (exchange.cuh): the exchange runs INSIDE the loss-finalize kernels over NVLink peer stores;
//     this file creates / exports / maps the mailboxes (CUDA IPC, one process per GPU) and offers the exchange as a
//     stand-alone one-warp kernel too (b200_allreduce_loss_peer / b200_allreduce_sums_peer);
//   * NCCL (b200_allreduce_loss / b200_allreduce_sums): ncclAllReduce on the caller's stream with a communicator the
//     caller creates through b200_nccl_comm_init.  libnccl is opened with dlopen at first use, so single-GPU users
//     need no NCCL at all.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"
#include "exchange.cuh"

// ---- peer mailboxes ------------------------------------------------------------------------------------
extern "C" size_t b200_peer_mailbox_bytes(void) { return B200_XCHG_MAILBOX_BYTES; }

extern "C" int b200_peer_mailbox_create(void** mailbox_out, void* ipc_handle_out) {
  B200_REQUIRE(mailbox_out, B200_ERR_BAD_ARG, "b200_peer_mailbox_create: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is exchanged as a 64-byte blob");
  void* p = nullptr;
  B200_CUDA(cudaMalloc(&p, B200_XCHG_MAILBOX_BYTES));  // plain cudaMalloc: exportable through CUDA IPC
  B200_CUDA(cudaMemset(p, 0, B200_XCHG_MAILBOX_BYTES));
  B200_CUDA(cudaDeviceSynchronize());
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      b200_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
      return B200_ERR_CUDA;
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
  }
  *mailbox_out = p;
  return B200_OK;
}

extern "C" int b200_peer_mailbox_open(const void* ipc_handle, void** mapped_out) {
  B200_REQUIRE(ipc_handle && mapped_out, B200_ERR_BAD_ARG, "b200_peer_mailbox_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  B200_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *mapped_out = p;
  return B200_OK;
}

extern "C" int b200_peer_mailbox_close(void* mapped) {
  if (mapped) B200_CUDA(cudaIpcCloseMemHandle(mapped));
  return B200_OK;
}

extern "C" int b200_peer_mailbox_destroy(void* mailbox) {
  if (mailbox) B200_CUDA(cudaFree(mailbox));
  return B200_OK;
}

extern "C" int b200_peer_mailbox_status(const void* own_mailbox, unsigned long long* epoch_out, unsigned int* errors_out) {
  B200_REQUIRE(own_mailbox, B200_ERR_BAD_ARG, "b200_peer_mailbox_status: null argument");
  XchgHeader h;
  B200_CUDA(cudaMemcpy(&h, own_mailbox, sizeof(h), cudaMemcpyDeviceToHost));
  if (epoch_out) *epoch_out = h.epoch;
  if (errors_out) *errors_out = h.errors;
  return B200_OK;
}

int b200_fill_exchange(B200Exchange& x, int rank, int world, void* const mailboxes[], const char* who) {
  x.rank = 0; x.world = 1;
  for (int r = 0; r < B200_XCHG_MAX_WORLD; ++r) x.mailbox[r] = nullptr;
  if (world <= 1) return B200_OK;
  B200_REQUIRE(world <= B200_XCHG_MAX_WORLD && rank >= 0 && rank < world && mailboxes, B200_ERR_BAD_ARG,
               "%s: bad exchange (rank %d of %d, at most %d ranks)", who, rank, world, B200_XCHG_MAX_WORLD);
  for (int r = 0; r < world; ++r) {
    B200_REQUIRE(mailboxes[r], B200_ERR_BAD_ARG, "%s: mailbox of rank %d is null", who, r);
    x.mailbox[r] = static_cast<unsigned char*>(mailboxes[r]);
  }
  x.rank = rank; x.world = world;
  return B200_OK;
}

template <typename T>
__global__ void __launch_bounds__(32) xchg_allreduce_kernel(B200Exchange x, T* values, int n) {
  const int lane = threadIdx.x;
  T v = (lane < n) ? values[lane] : (T)0;
  v = xchg_allreduce_warp<T>(x, v, n);
  if (lane < n) values[lane] = v;
}

template <typename T>
static int xchg_launch(T* values, int n, int rank, int world, void* const mailboxes[], void* stream, const char* who) {
  B200_REQUIRE(values && n >= 1 && n <= B200_XCHG_MAX_VALUES, B200_ERR_BAD_ARG, "%s: n must be in [1,%d]", who, B200_XCHG_MAX_VALUES);
  B200Exchange x;
  const int rc = b200_fill_exchange(x, rank, world, mailboxes, who);
  if (rc != B200_OK) return rc;
  if (world <= 1) return B200_OK;
  xchg_allreduce_kernel<T><<<1, 32, 0, (cudaStream_t)stream>>>(x, values, n);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// The collect half alone (the publish half ran inside a *_dp_publish call): values[n] <- sum over ranks; with loss_out the
// 12 YOLO terms are also folded in the reference's order (tyu:120-125).
__global__ void __launch_bounds__(32) xchg_collect_kernel(B200Exchange x, float* values, int n, float* loss_out) {
  const int lane = threadIdx.x;
  const float v = xchg_collect_warp<float>(x, n);
  if (lane < n && values) values[lane] = v;
  if (loss_out) {
    float total = 0.0f;
    for (int l = 0; l < 3; ++l) {
      const float t0 = __shfl_sync(0xffffffffu, v, l * 4 + 0), t1 = __shfl_sync(0xffffffffu, v, l * 4 + 1);
      const float t2 = __shfl_sync(0xffffffffu, v, l * 4 + 2), t3 = __shfl_sync(0xffffffffu, v, l * 4 + 3);
      total = __fadd_rn(total, __fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3));
    }
    if (lane == 0) *loss_out = total;
  }
}

extern "C" int b200_yolo_loss_collect_peer(float* out_parts, float* out_loss, int rank, int world, void* const mailboxes[], void* stream) {
  B200Exchange x;
  const int rc = b200_fill_exchange(x, rank, world, mailboxes, "b200_yolo_loss_collect_peer");
  if (rc != B200_OK) return rc;
  B200_REQUIRE(world > 1, B200_ERR_BAD_ARG, "b200_yolo_loss_collect_peer: nothing was published (world 1)");
  B200_REQUIRE(out_loss, B200_ERR_BAD_ARG, "b200_yolo_loss_collect_peer: null output");
  xchg_collect_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(x, out_parts, 12, out_loss);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_allreduce_loss_peer(float* partials, int n, int rank, int world, void* const mailboxes[], void* stream) {
  return xchg_launch<float>(partials, n, rank, world, mailboxes, stream, "b200_allreduce_loss_peer");
}
extern "C" int b200_allreduce_sums_peer(double* sums, int n, int rank, int world, void* const mailboxes[], void* stream) {
  return xchg_launch<double>(sums, n, rank, world, mailboxes, stream, "b200_allreduce_sums_peer");
}

// Protocol self-test on ONE device: `world` CTAs of one grid play the ranks (all co-resident, so the flag spins are
// safe), each with its own mailbox carved from `workspace` (world * b200_peer_mailbox_bytes(), zeroed by this call),
// for `rounds` consecutive exchanges of n values.  out[r][round][i] = the sum rank r obtained; every rank must
// see sum_q value(q, round, i) with value(q, round, i) = (q + 1) * (i + 1) + round.
__global__ void __launch_bounds__(32) xchg_selftest_kernel(unsigned char* ws, int world, int rounds, int n, float* out) {
  B200Exchange x;
  x.rank = blockIdx.x; x.world = world;
  for (int r = 0; r < B200_XCHG_MAX_WORLD; ++r) x.mailbox[r] = r < world ? ws + (size_t)r * B200_XCHG_MAILBOX_BYTES : nullptr;
  const int lane = threadIdx.x;
  for (int k = 0; k < rounds; ++k) {
    float v = lane < n ? (float)((x.rank + 1) * (lane + 1) + k) : 0.0f;
    v = xchg_allreduce_warp<float>(x, v, n);
    if (lane < n) out[((size_t)x.rank * rounds + k) * n + lane] = v;
  }
}

extern "C" int b200_peer_exchange_selftest(int world, int rounds, int n, void* workspace, size_t workspace_bytes, float* out,
                                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_REQUIRE(world >= 2 && world <= B200_XCHG_MAX_WORLD && rounds >= 1 && n >= 1 && n <= B200_XCHG_MAX_VALUES && workspace && out,
               B200_ERR_BAD_ARG, "b200_peer_exchange_selftest: bad argument");
  B200_REQUIRE(workspace_bytes >= (size_t)world * B200_XCHG_MAILBOX_BYTES, B200_ERR_WORKSPACE, "b200_peer_exchange_selftest: workspace too small");
  B200_CUDA(cudaMemsetAsync(workspace, 0, (size_t)world * B200_XCHG_MAILBOX_BYTES, stream));
  xchg_selftest_kernel<<<world, 32, 0, stream>>>(static_cast<unsigned char*>(workspace), world, rounds, n, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ---- NCCL transport ----------------------------------------------------------------------------------
// Minimal private declarations (ABI of NCCL 2.x; nccl.h is not needed to build).
typedef struct { char internal[128]; } b200_ncclUniqueId;
typedef int (*fn_ncclGetUniqueId)(b200_ncclUniqueId*);
typedef int (*fn_ncclCommInitRank)(void**, int, b200_ncclUniqueId, int);
typedef int (*fn_ncclCommDestroy)(void*);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_ncclGetErrorString)(int);
enum { B200_NCCL_FLOAT32 = 7, B200_NCCL_FLOAT64 = 8, B200_NCCL_SUM = 0 };

struct NcclApi {
  void* handle;
  fn_ncclGetUniqueId get_id; fn_ncclCommInitRank init_rank; fn_ncclCommDestroy destroy; fn_ncclAllReduce all_reduce;
  fn_ncclGetErrorString err;
};

static NcclApi* nccl_api() {
  static NcclApi api = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* env = getenv("B200_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.get_id = (fn_ncclGetUniqueId)dlsym(api.handle, "ncclGetUniqueId");
      api.init_rank = (fn_ncclCommInitRank)dlsym(api.handle, "ncclCommInitRank");
      api.destroy = (fn_ncclCommDestroy)dlsym(api.handle, "ncclCommDestroy");
      api.all_reduce = (fn_ncclAllReduce)dlsym(api.handle, "ncclAllReduce");
      api.err = (fn_ncclGetErrorString)dlsym(api.handle, "ncclGetErrorString");
    }
  }
  if (!api.handle || !api.get_id || !api.init_rank || !api.destroy || !api.all_reduce) return nullptr;
  return &api;
}

#define B200_NCCL(api, call, what)                                                              \
  do {                                                                                          \
    const int _r = (call);                                                                      \
    if (_r != 0) {                                                                              \
      b200_set_error("%s: NCCL error %d (%s)", what, _r, (api)->err ? (api)->err(_r) : "?");    \
      return B200_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

extern "C" int b200_nccl_unique_id(void* id128_out) {
  NcclApi* api = nccl_api();
  B200_REQUIRE(api, B200_ERR_UNSUPPORTED, "libnccl.so.2 could not be opened (set B200_NCCL_LIB to its path)");
  B200_REQUIRE(id128_out, B200_ERR_BAD_ARG, "b200_nccl_unique_id: null argument");
  b200_ncclUniqueId id;
  B200_NCCL(api, api->get_id(&id), "ncclGetUniqueId");
  memcpy(id128_out, &id, sizeof(id));
  return B200_OK;
}

extern "C" int b200_nccl_comm_init(void** comm_out, int world, int rank, const void* id128) {
  NcclApi* api = nccl_api();
  B200_REQUIRE(api, B200_ERR_UNSUPPORTED, "libnccl.so.2 could not be opened (set B200_NCCL_LIB to its path)");
  B200_REQUIRE(comm_out && id128 && world >= 1 && rank >= 0 && rank < world, B200_ERR_BAD_ARG, "b200_nccl_comm_init: bad argument");
  b200_ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  B200_NCCL(api, api->init_rank(&comm, world, id, rank), "ncclCommInitRank");
  *comm_out = comm;
  return B200_OK;
}

extern "C" int b200_nccl_comm_destroy(void* comm) {
  NcclApi* api = nccl_api();
  if (!api || !comm) return B200_OK;
  B200_NCCL(api, api->destroy(comm), "ncclCommDestroy");
  return B200_OK;
}

extern "C" int b200_allreduce_loss(void* comm, float* partials, int n, void* stream) {
  NcclApi* api = nccl_api();
  B200_REQUIRE(api, B200_ERR_UNSUPPORTED, "libnccl.so.2 could not be opened (set B200_NCCL_LIB to its path)");
  B200_REQUIRE(comm && partials && n >= 1, B200_ERR_BAD_ARG, "b200_allreduce_loss: bad argument");
  B200_NCCL(api, api->all_reduce(partials, partials, (size_t)n, B200_NCCL_FLOAT32, B200_NCCL_SUM, comm, (cudaStream_t)stream), "ncclAllReduce");
  return B200_OK;
}

extern "C" int b200_allreduce_sums(void* comm, double* sums, int n, void* stream) {
  NcclApi* api = nccl_api();
  B200_REQUIRE(api, B200_ERR_UNSUPPORTED, "libnccl.so.2 could not be opened (set B200_NCCL_LIB to its path)");
  B200_REQUIRE(comm && sums && n >= 1, B200_ERR_BAD_ARG, "b200_allreduce_sums: bad argument");
  B200_NCCL(api, api->all_reduce(sums, sums, (size_t)n, B200_NCCL_FLOAT64, B200_NCCL_SUM, comm, (cudaStream_t)stream), "ncclAllReduce");
  return B200_OK;
}
