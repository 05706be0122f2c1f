// NCCL plumbing for the sharded H_eff apply (one process per GPU).  libnccl is dlopen'ed so that the
// single-GPU library has no hard NCCL dependency; if torch already loaded its bundled libnccl.so.2 the
// same soname resolves to it.
#include <dlfcn.h>

#include <cstring>

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

void comm_destroy(Ctx* ctx) {
  if (ctx->nccl_comm) {
    ctx->sync();
    nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->world = 1;
  }
}

void comm_allreduce_sum(Ctx* ctx, double* buf, int64_t n) {
  Ctx::Scope prof_scope(ctx, 3);
  prof_scope.r.tiles = n <= 64 ? 0 : 3;            // kind (tnl_profile_collectives)
  TNL_CHECK(ctx->nccl_comm, "communicator not initialised");
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
