// NCCL plumbing for the sharded H_eff apply (one process per GPU).  libnccl is dlopen'ed so that the
// single-GPU library has no hard NCCL dependency; if torch already loaded its bundled libnccl.so.2 the
// same soname resolves to it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "env.hpp"

namespace tnl {

// minimal NCCL ABI (nccl.h, stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclFloat64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl() {
  static NcclApi api;
  if (!api.lib) {
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) throw Error(5, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) {
      void* p = dlsym(api.lib, n);
      if (!p) throw Error(5, std::string("libnccl lacks symbol ") + n);
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.ReduceScatter = (decltype(api.ReduceScatter))sym("ncclReduceScatter");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  }
  return api;
}

#define NCCL_OK(call)                                                                                   \
  do {                                                                                                  \
    ncclResult_t r__ = (call);                                                                          \
    if (r__ != 0) throw ::tnl::Error(5, std::string("NCCL error: ") + nccl().GetErrorString(r__) + " in " #call); \
  } while (0)

void comm_unique_id(char* out128) {
  ncclUniqueId id;
  NCCL_OK(nccl().GetUniqueId(&id));
  std::memcpy(out128, id.internal, 128);
}

void comm_init(Ctx* ctx, const char* uid128, int rank, int world) {
  TNL_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank / world size");
  ncclUniqueId id;
  std::memcpy(id.internal, uid128, 128);
  ncclComm_t comm;
  CUDA_OK(cudaSetDevice(ctx->device));
  NCCL_OK(nccl().CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
}

static void stage_release(Ctx* ctx) {
  Ctx::PeerStage& ps = ctx->pstage;
  for (int k = 0; k < (int)ps.peer_base.size(); k++)
    if (k != ctx->rank && ps.peer_base[k]) cudaIpcCloseMemHandle(ps.peer_base[k]);
  ps.peer_base.clear();
  if (ps.base) cudaFree(ps.base);
  if (ps.d_peer_slots) cudaFree(ps.d_peer_slots);
  if (ps.d_peer_flags) cudaFree(ps.d_peer_flags);
  if (ps.d_done) cudaFree(ps.d_done);
  ps.base = nullptr; ps.d_peer_slots = nullptr; ps.d_peer_flags = nullptr; ps.d_done = nullptr;
  ps.slot_cap = 0;
  ps.ok = false;
}

// Collective: (re)allocates the staging area when a slot of `slot_doubles` does not fit, exchanges the CUDA-IPC handles
// through the NCCL communicator and maps the peers' areas.  Every rank takes the same decisions (sizes derive from
// the replicated plan), and a failure on any rank switches the fused path off on all of them.
// header of a staging area (bytes): [0, 512) "ready" words, [512, 1024) "consumed" words (fused reduce-scatter);
// [1024, 2048) sequence words [2 parities][64 sources] and [2048, 10240) mailboxes [2][64][8 doubles] of the scalar
// all-reduce; slots start at 16 KB
constexpr size_t kStageFlagBytes = 16384;
bool comm_stage_ensure(Ctx* ctx, size_t slot_doubles) {
  Ctx::PeerStage& ps = ctx->pstage;
  if (ctx->world <= 1 || !ctx->nccl_comm) return false;
  if (ps.tried && !ps.ok) return false;
  if (ps.ok && slot_doubles <= ps.slot_cap) return true;
  if (const char* e = getenv("TNL_FUSED_RS")) if (atoi(e) == 0) { ps.tried = true; ps.ok = false; return false; }
  if (const char* e = getenv("TNL_PEER_SCALAR_AR")) ps.small_ar = atoi(e) != 0;
  TNL_CHECK(ctx->world <= 64, "staging flags are laid out for at most 64 ranks");
  ps.tried = true;
  ctx->sync();
  // quiesce: nobody may still be writing into an area that is about to be unmapped
  double* dflag = (double*)ctx->alloc(8 * sizeof(double));
  CUDA_OK(cudaMemsetAsync(dflag, 0, 8 * sizeof(double), ctx->stream));
  comm_allreduce_sum(ctx, dflag, 1);
  ctx->sync();
  stage_release(ctx);
  const size_t cap = (std::max<size_t>(slot_doubles + slot_doubles / 2, size_t(1) << 20) + 1) & ~size_t(1);
  const size_t bytes = kStageFlagBytes + cap * ctx->world * sizeof(double);
  int good = 1;
  cudaIpcMemHandle_t mine{};
  if (cudaMalloc(&ps.base, bytes) != cudaSuccess) { good = 0; ps.base = nullptr; cudaGetLastError(); }
  if (good) {
    CUDA_OK(cudaMemset(ps.base, 0, bytes));
    if (cudaIpcGetMemHandle(&mine, ps.base) != cudaSuccess) { good = 0; cudaGetLastError(); }
  }
  // all-gather of the 64-byte handles (+ one word "good") through NCCL: 9 doubles per rank
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  std::vector<double> sendh(9, 0.0), recvh((size_t)9 * ctx->world, 0.0);
  std::memcpy(sendh.data(), &mine, 64);
  sendh[8] = good ? 1.0 : 0.0;
  double* dsend = (double*)ctx->alloc(9 * sizeof(double));
  double* drecv = (double*)ctx->alloc((size_t)9 * ctx->world * sizeof(double));
  CUDA_OK(cudaMemcpyAsync(dsend, sendh.data(), 9 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  comm_allgather(ctx, dsend, drecv, 9);
  CUDA_OK(cudaMemcpyAsync(recvh.data(), drecv, recvh.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  for (int k = 0; k < ctx->world; k++) good = good && recvh[(size_t)9 * k + 8] == 1.0;
  ps.peer_base.assign(ctx->world, nullptr);
  if (good) {
    for (int k = 0; k < ctx->world && good; k++) {
      if (k == ctx->rank) { ps.peer_base[k] = ps.base; continue; }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, &recvh[(size_t)9 * k], 64);
      if (cudaIpcOpenMemHandle(&ps.peer_base[k], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = 0; ps.peer_base[k] = nullptr; cudaGetLastError(); }
    }
  }
  // agree on the outcome
  double g = good ? 0.0 : 1.0;
  CUDA_OK(cudaMemcpyAsync(dflag, &g, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  comm_allreduce_sum(ctx, dflag, 1);
  CUDA_OK(cudaMemcpyAsync(&g, dflag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  ctx->free(dflag); ctx->free(dsend); ctx->free(drecv);
  if (g != 0.0) { stage_release(ctx); return false; }
  ps.slot_cap = cap;
  std::vector<double*> slots(ctx->world);
  std::vector<unsigned long long*> flags(ctx->world);
  for (int k = 0; k < ctx->world; k++) {
    flags[k] = (unsigned long long*)ps.peer_base[k];
    slots[k] = (double*)((char*)ps.peer_base[k] + kStageFlagBytes) + (size_t)ctx->rank * cap;   // MY slot on rank k
  }
  CUDA_OK(cudaMalloc(&ps.d_peer_slots, ctx->world * sizeof(double*)));
  CUDA_OK(cudaMalloc(&ps.d_peer_flags, ctx->world * sizeof(unsigned long long*)));
  CUDA_OK(cudaMalloc(&ps.d_done, sizeof(unsigned int)));
  CUDA_OK(cudaMemcpy(ps.d_peer_slots, slots.data(), ctx->world * sizeof(double*), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(ps.d_peer_flags, flags.data(), ctx->world * sizeof(unsigned long long*), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemset(ps.d_done, 0, sizeof(unsigned int)));
  ps.epoch = 0;                        // flags were zeroed with the new area
  ps.ar_epoch = 0;
  ps.ok = true;
  return true;
}

void comm_destroy(Ctx* ctx) {
  if (ctx->nccl_comm) {
    ctx->sync();
    stage_release(ctx);
    ctx->pstage.tried = false;
    nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->world = 1;
  }
}

void peer_allreduce_small(Ctx* ctx, double* buf, int n);   // kernels.cu

void comm_allreduce_sum(Ctx* ctx, double* buf, int64_t n) {
  Ctx::Scope prof_scope(ctx, 3);
  prof_scope.r.tiles = n <= 64 ? 0 : 3;            // kind (tnl_profile_collectives)
  TNL_CHECK(ctx->nccl_comm, "communicator not initialised");
  if (n <= 8 && ctx->pstage.ok && ctx->pstage.small_ar) {
    // Krylov inner products: a few doubles, latency only -- mailboxes in the peers' staging headers instead of an
    // NCCL kernel (every rank sums the W contributions in rank order: identical result everywhere)
    peer_allreduce_small(ctx, buf, (int)n);
    ctx->cnt.allreduce_bytes += 8.0 * n;
    return;
  }
  NCCL_OK(nccl().AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  ctx->cnt.allreduce_bytes += 8.0 * n;
}

// recv[0..n) = sum over ranks of send[rank*n .. rank*n+n)
void comm_reduce_scatter_sum(Ctx* ctx, const double* send, double* recv, int64_t n) {
  Ctx::Scope prof_scope(ctx, 3);
  prof_scope.r.tiles = 1;
  TNL_CHECK(ctx->nccl_comm, "communicator not initialised");
  NCCL_OK(nccl().ReduceScatter(send, recv, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  ctx->cnt.allreduce_bytes += 8.0 * n * ctx->world;
}
// recv[k*n .. k*n+n) = send of rank k
void comm_allgather(Ctx* ctx, const double* send, double* recv, int64_t n) {
  Ctx::Scope prof_scope(ctx, 3);
  prof_scope.r.tiles = 2;
  TNL_CHECK(ctx->nccl_comm, "communicator not initialised");
  NCCL_OK(nccl().AllGather(send, recv, (size_t)n, ncclFloat64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  ctx->cnt.allreduce_bytes += 8.0 * n * ctx->world;
}

// contiguous share of a sector of dimension d for `rank` out of `world`; remainders rotate with the sector number
void shard_range(int d, int world, int sector, int rank, int* start, int* count) {
  int base = d / world, rem = d % world;
  int rr = (rank + sector) % world;
  *count = base + (rr < rem ? 1 : 0);
  *start = rr * base + std::min(rr, rem);
}

}  // namespace tnl
