// extern "C" boundary of libtnl_b200.so -- see include/tnl_b200.h for the contract and the
// reference interfaces each entry point replaces.
#include "../../include/tnl_b200.h"

#include <cmath>
#include <cstring>

#include "algebra.hpp"
#include "env.hpp"
#include "krylov.hpp"

using namespace tnl;

struct tnl_ctx_s { Ctx ctx; cudaEvent_t ev[8] = {}; double cat_ms[8] = {}; double coll_ms[12] = {}; explicit tnl_ctx_s(int dev) : ctx(dev) {} };
struct tnl_tensor_s { TensorP t; };
struct tnl_env_s { Env env; tnl_env_s(Ctx* c, int n) : env(c, n) {} };
struct tnl_sumop_s { SumOp op; explicit tnl_sumop_s(Ctx* c) : op(c) {} };

static thread_local std::string g_last_error;

template <class F>
static int guard(Ctx* ctx, F&& f) {
  try {
    f();
    return 0;
  } catch (const Error& e) {
    g_last_error = e.what();
    if (ctx) ctx->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    if (ctx) ctx->last_error = e.what();
    return 1;
  }
}

static std::vector<Index> make_inds(int rank, int nq, const tnl_index_t* inds) {
  TNL_CHECK(rank >= 1 && rank <= MAXR, "rank out of range");
  TNL_CHECK(nq >= 1 && nq <= MAXQ, "number of conserved charges out of range");
  std::vector<Index> out(rank);
  for (int k = 0; k < rank; k++) {
    TNL_CHECK(inds[k].nsect >= 1 && inds[k].nsect < 1024, "sector count out of range");
    TNL_CHECK(inds[k].dir == 1 || inds[k].dir == -1, "arrow must be +1 or -1");
    out[k].nq = nq;
    out[k].dir = inds[k].dir;
    for (int s = 0; s < inds[k].nsect; s++) {
      TNL_CHECK(inds[k].dims[s] >= 1, "sector dims must be positive");
      out[k].dims.push_back(inds[k].dims[s]);
      Charge q = charge_zero();
      for (int a = 0; a < nq; a++) q[a] = inds[k].qns[s * nq + a];
      out[k].qns.push_back(q);
    }
  }
  return out;
}

static HostBlocks make_host(int rank, int nq, const tnl_index_t* inds, int64_t nblocks, const int32_t* coords,
                            const int64_t* offsets, const double* data) {
  HostBlocks hb;
  hb.rank = rank;
  hb.inds = make_inds(rank, nq, inds);
  for (int64_t n = 0; n < nblocks; n++) {
    std::vector<int> c(rank);
    for (int k = 0; k < rank; k++) c[k] = coords[n * rank + k];
    hb.coords.push_back(c);
    hb.offsets.push_back(offsets[n]);
  }
  hb.data = data;
  return hb;
}

extern "C" {

int tnl_ctx_create(int device, tnl_ctx_t* out) {
  return guard(nullptr, [&] {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(3, "tnl_b200 needs a CUDA device: no GPU visible (there is no CPU fallback)");
    *out = new tnl_ctx_s(device);
  });
}
int tnl_ctx_destroy(tnl_ctx_t c) {
  return guard(nullptr, [&] {
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    delete c;
  });
}
const char* tnl_last_error(tnl_ctx_t c) { return c ? c->ctx.last_error.c_str() : g_last_error.c_str(); }
int tnl_get_counters(tnl_ctx_t c, double* o) {
  return guard(&c->ctx, [&] {
    const Counters& k = c->ctx.cnt;
    o[0] = k.gemm_flops; o[1] = k.xf_flops; o[2] = k.vec_bytes; o[3] = k.xf_bytes;
    o[4] = (double)k.launches; o[5] = (double)k.gemm_launches; o[6] = k.apply_count; o[7] = k.allreduce_bytes;
  });
}
int tnl_reset_counters(tnl_ctx_t c) { return guard(&c->ctx, [&] { c->ctx.cnt = Counters(); }); }
int tnl_ctx_reserve(tnl_ctx_t c, int64_t bytes) {
  return guard(&c->ctx, [&] {
    // grow the stream-ordered pool once (release threshold = never): later allocations are carved from it
    // without asking the driver for new physical memory in the middle of a sweep
    void* p = nullptr;
    CUDA_OK(cudaMallocAsync(&p, (size_t)bytes, c->ctx.stream));
    CUDA_OK(cudaFreeAsync(p, c->ctx.stream));
    c->ctx.sync();
  });
}
int tnl_ctx_sync(tnl_ctx_t c) { return guard(&c->ctx, [&] { c->ctx.sync(); }); }
int tnl_timer_start(tnl_ctx_t c, int32_t slot) {
  return guard(&c->ctx, [&] {
    TNL_CHECK(slot >= 0 && slot < 4, "timer slot out of range");
    if (!c->ev[2 * slot]) { CUDA_OK(cudaEventCreate(&c->ev[2 * slot])); CUDA_OK(cudaEventCreate(&c->ev[2 * slot + 1])); }
    CUDA_OK(cudaEventRecord(c->ev[2 * slot], c->ctx.stream));
  });
}
int tnl_timer_stop(tnl_ctx_t c, int32_t slot, double* ms) {
  return guard(&c->ctx, [&] {
    TNL_CHECK(slot >= 0 && slot < 4 && c->ev[2 * slot], "timer not started");
    CUDA_OK(cudaEventRecord(c->ev[2 * slot + 1], c->ctx.stream));
    CUDA_OK(cudaEventSynchronize(c->ev[2 * slot + 1]));
    float f = 0;
    CUDA_OK(cudaEventElapsedTime(&f, c->ev[2 * slot], c->ev[2 * slot + 1]));
    *ms = f;
  });
}

int tnl_comm_unique_id(char* out128) {
  return guard(nullptr, [&] { comm_unique_id(out128); });
}
int tnl_comm_init(tnl_ctx_t c, const char* uid128, int32_t rank, int32_t world) {
  return guard(&c->ctx, [&] { comm_init(&c->ctx, uid128, rank, world); });
}
int tnl_comm_destroy(tnl_ctx_t c) {
  return guard(&c->ctx, [&] { comm_destroy(&c->ctx); });
}
int tnl_comm_set_sharding(tnl_ctx_t c, int32_t enable) {
  return guard(&c->ctx, [&] {
    c->ctx.sync();
    c->ctx.shard_enabled = enable != 0;
    c->ctx.slot_epoch++;                       // every cached apply plan is rebuilt for the new mode
  });
}
int tnl_comm_bench(tnl_ctx_t c, int64_t n, int32_t reps, int32_t kind, double* ms) {
  return guard(&c->ctx, [&] {
    Ctx* ctx = &c->ctx;
    const int W = ctx->world;
    double* a = (double*)ctx->alloc((size_t)n * W * sizeof(double));
    double* b = (double*)ctx->alloc((size_t)n * W * sizeof(double));
    CUDA_OK(cudaMemsetAsync(a, 0, (size_t)n * W * sizeof(double), ctx->stream));
    auto go = [&] {
      if (kind == 0) comm_allreduce_sum(ctx, a, n);
      else if (kind == 1) comm_reduce_scatter_sum(ctx, a, b, n);
      else comm_allgather(ctx, a, b, n);
    };
    for (int i = 0; i < 3; i++) go();
    ctx->sync();
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    CUDA_OK(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < reps; i++) go();
    CUDA_OK(cudaEventRecord(e1, ctx->stream));
    CUDA_OK(cudaEventSynchronize(e1));
    float f = 0;
    CUDA_OK(cudaEventElapsedTime(&f, e0, e1));
    *ms = f / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->free(a); ctx->free(b);
  });
}
int tnl_shard_range(int32_t dim, int32_t world, int32_t sector, int32_t rank, int32_t* start, int32_t* count) {
  return guard(nullptr, [&] {
    TNL_CHECK(world >= 1 && rank >= 0 && rank < world && dim >= 0, "bad shard query");
    int s, n;
    shard_range(dim, world, sector, rank, &s, &n);
    *start = s; *count = n;
  });
}
int tnl_gemm_selftest(tnl_ctx_t c, int32_t M, int32_t N, int32_t K, int32_t transA, int32_t transB, int32_t reps,
                      int32_t verify, double* ms, double* maxerr) {
  return guard(&c->ctx, [&] { gemm_selftest(&c->ctx, M, N, K, transA != 0, transB != 0, reps, verify != 0, ms, maxerr); });
}
int tnl_profile_gemm(tnl_ctx_t c, int32_t enable) {
  return guard(&c->ctx, [&] { c->ctx.prof_gemm = enable != 0; });
}
int tnl_profile_read(tnl_ctx_t c, double* total_ms, int64_t* launches, double* flops, double* max_tflops) {
  return guard(&c->ctx, [&] {
    c->ctx.sync();
    *total_ms = 0; *launches = 0; *flops = 0; *max_tflops = 0;
    for (int k = 0; k < 8; k++) c->cat_ms[k] = 0;
    for (int k = 0; k < 12; k++) c->coll_ms[k] = 0;
    for (auto& r : c->ctx.prof_recs) {
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
      if (r.cat == 0) {
        *total_ms += ms; *launches += 1; *flops += r.flops;
        if (ms > 0 && r.flops > 1e9) *max_tflops = std::max(*max_tflops, r.flops / (ms * 1e-3) / 1e12);
      }
      if (r.cat >= 0 && r.cat < 4) { c->cat_ms[r.cat] += ms; c->cat_ms[4 + r.cat] += 1; }
      if (r.cat == 3 && r.tiles >= 0 && r.tiles < 6) { c->coll_ms[r.tiles] += ms; c->coll_ms[6 + r.tiles] += 1; }
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    c->ctx.prof_recs.clear();
  });
}
/* out[0..3] = device ms of the last tnl_profile_read by category (gemm, transform, vector, collective),
 * out[4..7] = launch counts */
int tnl_profile_categories(tnl_ctx_t c, double* out8) {
  return guard(&c->ctx, [&] { for (int k = 0; k < 8; k++) out8[k] = c->cat_ms[k]; out8[7] = c->ctx.cnt.host_plan_ms; });
}

/* collectives of the last tnl_profile_read by kind: out[0..5] = ms of scalar all-reduces (inner products), NCCL
 * reduce-scatters (H_eff partial sums), all-gathers, large all-reduces (truncation factors), and of the fused path the
 * "slots consumed" wait and the ordered sum over the staging slots (includes the wait for the peers' data);
 * out[6..11] = call counts */
int tnl_profile_collectives(tnl_ctx_t c, double* out12) {
  return guard(&c->ctx, [&] { for (int k = 0; k < 12; k++) out12[k] = c->coll_ms[k]; });
}

int tnl_tensor_import(tnl_ctx_t c, int32_t rank, int32_t nq, const tnl_index_t* inds, int64_t nblocks,
                      const int32_t* coords, const int64_t* offsets, const double* data, int32_t nrow,
                      tnl_tensor_t* out) {
  return guard(&c->ctx, [&] {
    HostBlocks hb = make_host(rank, nq, inds, nblocks, coords, offsets, data);
    *out = new tnl_tensor_s{import_tensor(&c->ctx, hb, nrow)};
  });
}
int tnl_tensor_import_c128(tnl_ctx_t c, int32_t rank, int32_t nq, const tnl_index_t* inds, int64_t nblocks,
                           const int32_t* coords, const int64_t* offsets, const double* data_re_im, int32_t nrow,
                           tnl_tensor_t* out) {
  return guard(&c->ctx, [&] {
    HostBlocks hb = make_host(rank, nq, inds, nblocks, coords, offsets, data_re_im);
    hb.cplx = true;
    *out = new tnl_tensor_s{import_tensor(&c->ctx, hb, nrow)};
  });
}
int tnl_tensor_is_complex(tnl_tensor_t t, int32_t* out) {
  return guard(t->t->ctx, [&] { *out = t->t->cplx ? 1 : 0; });
}
// real -> planar complex with a zero imaginary plane (no-op on complex tensors)
static TensorP promoted(const TensorP& t) {
  if (t->cplx) return t;
  auto n = std::make_shared<Tensor>(t->ctx, t->inds, t->nrow, true, true);
  vec_copy(t->ctx, n->d, t->d, t->nelem);
  n->present = t->present;
  return n;
}
int tnl_tensor_promote(tnl_tensor_t t) {
  return guard(t->t->ctx, [&] { t->t = promoted(t->t); });
}
int tnl_tensor_create(tnl_ctx_t c, int32_t rank, int32_t nq, const tnl_index_t* inds, int32_t nrow, tnl_tensor_t* out) {
  return guard(&c->ctx, [&] {
    *out = new tnl_tensor_s{std::make_shared<Tensor>(&c->ctx, make_inds(rank, nq, inds), nrow)};
  });
}
int tnl_tensor_free(tnl_tensor_t t) {
  return guard(nullptr, [&] { delete t; });
}
int tnl_tensor_copy(tnl_tensor_t t, tnl_tensor_t* out) {
  return guard(t->t->ctx, [&] {
    auto n = std::make_shared<Tensor>(t->t->ctx, t->t->inds, t->t->nrow, true, t->t->cplx);
    vec_copy(t->t->ctx, n->d, t->t->d, t->t->planes() * t->t->nelem);
    n->present = t->t->present;
    *out = new tnl_tensor_s{n};
  });
}
int tnl_tensor_rank(tnl_tensor_t t, int32_t* rank, int32_t* nq) {
  return guard(t->t->ctx, [&] { *rank = t->t->rank(); *nq = t->t->inds[0].nq; });
}
int tnl_tensor_nrow(tnl_tensor_t t, int32_t* nrow) {
  return guard(t->t->ctx, [&] { *nrow = t->t->nrow; });
}
int tnl_tensor_index(tnl_tensor_t t, int32_t which, int32_t* nsect, int32_t* dir, int32_t* dims, int32_t* qns, int32_t cap) {
  return guard(t->t->ctx, [&] {
    TNL_CHECK(which >= 0 && which < t->t->rank(), "index number out of range");
    const Index& ix = t->t->inds[which];
    *nsect = ix.nsect();
    *dir = ix.dir;
    if (dims && qns) {
      TNL_CHECK(cap >= ix.nsect(), "index buffer too small");
      for (int s = 0; s < ix.nsect(); s++) {
        dims[s] = ix.dims[s];
        for (int a = 0; a < ix.nq; a++) qns[s * ix.nq + a] = ix.qns[s][a];
      }
    }
  });
}
int tnl_tensor_export_size(tnl_tensor_t t, int64_t* nblocks, int64_t* nelem) {
  return guard(t->t->ctx, [&] {
    *nblocks = (int64_t)t->t->blocks.size();
    *nelem = t->t->logical_elems();
  });
}
int tnl_tensor_export(tnl_tensor_t t, int32_t* coords, int64_t* offsets, double* data) {
  return guard(t->t->ctx, [&] {
    Ctx* ctx = t->t->ctx;
    TensorP nat = to_natural(ctx, *t->t);      // contiguous column-major blocks, column-major block order
    const int r = nat->rank();
    for (size_t b = 0; b < nat->blocks.size(); b++) {
      for (int k = 0; k < r; k++) coords[b * r + k] = nat->blocks[b].c[k];
      offsets[b] = nat->blocks[b].off;
    }
    int64_t n = nat->logical_elems();
    if (!nat->cplx) {
      if (n) CUDA_OK(cudaMemcpyAsync(data, nat->d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
    } else {
      // complex tensors come back interleaved (re, im) like NDTensors ComplexF64 storage: 2 * nelem doubles
      std::vector<double> re((size_t)n), im((size_t)n);
      if (n) {
        CUDA_OK(cudaMemcpyAsync(re.data(), nat->d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_OK(cudaMemcpyAsync(im.data(), nat->im(), n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      }
      ctx->sync();
      for (int64_t e = 0; e < n; e++) { data[2 * e] = re[e]; data[2 * e + 1] = im[e]; }
    }
  });
}
int tnl_tensor_scale_index(tnl_tensor_t t, int32_t which, const double* values) {
  return guard(t->t->ctx, [&] {
    Ctx* ctx = t->t->ctx;
    TNL_CHECK(which >= 0 && which < t->t->rank(), "index number out of range");
    std::vector<double> w(values, values + t->t->inds[which].dim());
    double* d = ctx->upload(w);
    scale_index(ctx, *t->t, which, d);
    ctx->sync();
    ctx->free(d);
  });
}
int tnl_tensor_fill_random(tnl_tensor_t t, uint64_t seed) {
  return guard(t->t->ctx, [&] {
    fill_random(t->t->ctx, *t->t, seed);
    if (t->t->cplx) {                          // second plane from an independent counter stream
      Tensor v(t->t->ctx, t->t->inds, t->t->nrow, false);
      v.d = t->t->im();
      v.owns = false;
      fill_random(t->t->ctx, v, seed ^ 0x5DEECE66Dull);
    }
    t->t->ctx->sync();
  });
}

static void same_layout(const Tensor& a, const Tensor& b) {
  TNL_CHECK(a.nelem == b.nelem && a.nrow == b.nrow && a.rank() == b.rank() && a.cplx == b.cplx,
            "vectors have different layouts");
}
int tnl_vec_dot(tnl_tensor_t x, tnl_tensor_t y, double* out) {
  return guard(x->t->ctx, [&] {
    same_layout(*x->t, *y->t);
    Ctx* ctx = x->t->ctx;
    vec_dot(ctx, x->t->d, y->t->d, x->t->planes() * x->t->nelem, 0);   // complex: the real part of <x, y>
    fetch_scalars(ctx, 1);
    *out = ctx->h_scalars[0];
  });
}
int tnl_vec_dot_c(tnl_tensor_t x, tnl_tensor_t y, double* re, double* im) {
  return guard(x->t->ctx, [&] {
    same_layout(*x->t, *y->t);
    Ctx* ctx = x->t->ctx;
    if (x->t->cplx) {
      vec_cdot(ctx, x->t->d, y->t->d, x->t->nelem, 0);
      fetch_scalars(ctx, 2);
      *re = ctx->h_scalars[0]; *im = ctx->h_scalars[1];
    } else {
      vec_dot(ctx, x->t->d, y->t->d, x->t->nelem, 0);
      fetch_scalars(ctx, 1);
      *re = ctx->h_scalars[0]; *im = 0.0;
    }
  });
}
int tnl_vec_norm(tnl_tensor_t x, double* out) {
  double d = 0;
  int rc = tnl_vec_dot(x, x, &d);
  *out = std::sqrt(d);
  return rc;
}
int tnl_vec_scale(tnl_tensor_t x, double a) {
  return guard(x->t->ctx, [&] { vec_scale(x->t->ctx, x->t->d, x->t->planes() * x->t->nelem, a); });
}
int tnl_vec_axpy(tnl_tensor_t y, tnl_tensor_t x, double a) {
  return guard(x->t->ctx, [&] {
    same_layout(*x->t, *y->t);
    vec_axpy(x->t->ctx, y->t->d, x->t->d, x->t->planes() * x->t->nelem, a);
  });
}

int tnl_env_create(tnl_ctx_t c, int32_t nsites, tnl_env_t* env) {
  return guard(&c->ctx, [&] {
    TNL_CHECK(nsites >= 2, "need at least two sites");
    *env = new tnl_env_s(&c->ctx, nsites);
  });
}
int tnl_env_destroy(tnl_env_t e) {
  return guard(nullptr, [&] { delete e; });
}
int tnl_env_set_site_op(tnl_env_t e, int32_t site, int32_t nq, const tnl_index_t* inds4, int64_t nblocks,
                        const int32_t* coords, const int64_t* offsets, const double* data) {
  return guard(e->env.ctx, [&] {
    HostBlocks hb = make_host(4, nq, inds4, nblocks, coords, offsets, data);
    e->env.set_site_op(site, hb);
  });
}
int tnl_env_set_site_op_term(tnl_env_t e, int32_t term, int32_t site, int32_t nq, const tnl_index_t* inds4,
                             int64_t nblocks, const int32_t* coords, const int64_t* offsets, const double* data) {
  return guard(e->env.ctx, [&] {
    HostBlocks hb = make_host(4, nq, inds4, nblocks, coords, offsets, data);
    e->env.term(term).set_site_op(site, hb);
  });
}
int tnl_env_update_site_op(tnl_env_t e, int32_t site, int32_t nq, const tnl_index_t* inds4, int64_t nblocks,
                           const int32_t* coords, const int64_t* offsets, const double* data) {
  return guard(e->env.ctx, [&] {
    HostBlocks hb = make_host(4, nq, inds4, nblocks, coords, offsets, data);
    e->env.update_site_op(site, hb);
  });
}
int tnl_env_cm_set_term(tnl_env_t e, int32_t site, int64_t id, int32_t has_wl, int32_t has_wr, int32_t nq,
                        const tnl_index_t* inds4, int64_t nblocks, const int32_t* coords, const int64_t* offsets,
                        const double* data) {
  return guard(e->env.ctx, [&] {
    HostBlocks hb = make_host(4, nq, inds4, nblocks, coords, offsets, data);
    e->env.cm_set_term(site, id, hb, has_wl != 0, has_wr != 0);
  });
}
int tnl_env_add_penalty(tnl_env_t e, double weight, int32_t nsites, const tnl_tensor_t* tensors) {
  return guard(e->env.ctx, [&] {
    std::vector<TensorP> M;
    for (int j = 0; j < nsites; j++) {
      TNL_CHECK(tensors[j] && tensors[j]->t->rank() == 3, "penalised MPS tensors must be (l, s, r)");
      M.push_back(tensors[j]->t);
    }
    e->env.add_penalty(M, weight);
  });
}
int tnl_env_set_state(tnl_env_t e, int32_t site, tnl_tensor_t a) {
  return guard(e->env.ctx, [&] { e->env.set_state(site, a->t); });
}
int tnl_env_get_state(tnl_env_t e, int32_t site, tnl_tensor_t* out) {
  return guard(e->env.ctx, [&] {
    TNL_CHECK(site >= 1 && site <= e->env.N && e->env.A[site - 1], "site out of range or unset");
    *out = new tnl_tensor_s{e->env.A[site - 1]};
  });
}
int tnl_env_set_nsite(tnl_env_t e, int32_t nsite) {
  return guard(e->env.ctx, [&] {
    TNL_CHECK(nsite >= 0 && nsite <= 2, "nsite must be 0, 1 or 2");
    e->env.set_nsite(nsite);
  });
}
int tnl_env_position(tnl_env_t e, int32_t pos) {
  return guard(e->env.ctx, [&] { e->env.position(pos); });
}
int tnl_env_move_center(tnl_env_t e, int32_t from, int32_t to) {
  return guard(e->env.ctx, [&] { e->env.move_center(from, to); });
}
int tnl_env_make_phi(tnl_env_t e, int32_t pos, tnl_tensor_t* phi) {
  return guard(e->env.ctx, [&] { *phi = new tnl_tensor_s{e->env.make_phi(pos)}; });
}
int tnl_env_apply_flops(tnl_env_t e, double* flops) {
  return guard(e->env.ctx, [&] { *flops = e->env.apply_flops(); });
}
int tnl_heff_apply(tnl_env_t e, tnl_tensor_t v, tnl_tensor_t* out) {
  return guard(e->env.ctx, [&] {
    TensorP vq = v->t->nrow == 1 ? v->t : relayout(e->env.ctx, *v->t, 1);
    if (e->env.complex_at_position()) vq = promoted(vq);       // complex environments: complex product
    auto o = std::make_shared<Tensor>(e->env.ctx, vq->inds, 1, true, vq->cplx);
    e->env.apply(*vq, *o);
    e->env.ctx->sync();
    *out = new tnl_tensor_s{o};
  });
}
int tnl_eigsolve_lanczos(tnl_env_t e, tnl_tensor_t phi, double tol, int32_t krylovdim, int32_t maxiter, int32_t eager,
                         double* eval, int32_t* converged, int32_t* numops, int32_t* numiter, double* normres) {
  return guard(e->env.ctx, [&] {
    if (phi->t->nrow != 1) phi->t = relayout(e->env.ctx, *phi->t, 1);
    if (e->env.complex_at_position()) phi->t = promoted(phi->t);   // complex environments: complex Ritz vector
    LanczosResult r = e->env.eigsolve(*phi->t, tol, krylovdim, maxiter, eager != 0);
    *eval = r.eval; *converged = r.converged; *numops = r.numops; *numiter = r.numiter; *normres = r.normres;
  });
}
int tnl_exponentiate(tnl_env_t e, tnl_tensor_t phi, double t_re, double t_im, double tol, int32_t krylovdim,
                     int32_t maxiter, int32_t eager, int32_t* converged, int32_t* numops, int32_t* numiter, double* err) {
  return guard(e->env.ctx, [&] {
    if (phi->t->nrow != 1) phi->t = relayout(e->env.ctx, *phi->t, 1);
    if (t_im != 0.0 || e->env.complex_at_position()) phi->t = promoted(phi->t);   // real-time step: ComplexF64 from here on
    ExpResult r = e->env.exponentiate(*phi->t, t_re, t_im, tol, krylovdim, maxiter, eager != 0);
    *converged = r.converged; *numops = r.numops; *numiter = r.numiter; *err = r.err;
  });
}
int tnl_env_absorb_bond(tnl_env_t e, int32_t pos, int32_t ortho_left, tnl_tensor_t carry) {
  return guard(e->env.ctx, [&] { e->env.absorb_bond(pos, ortho_left != 0, *carry->t); });
}
int tnl_expectation(tnl_env_t e, tnl_tensor_t phi, double* out) {
  return guard(e->env.ctx, [&] {
    TensorP vq = phi->t->nrow == 1 ? phi->t : relayout(e->env.ctx, *phi->t, 1);
    if (e->env.complex_at_position()) vq = promoted(vq);
    *out = e->env.expectation(*vq);
  });
}
int tnl_replacebond(tnl_env_t e, int32_t pos, tnl_tensor_t phi, int32_t ortho_left, int64_t maxdim, int64_t mindim,
                    double cutoff, double noise, int32_t normalize, int32_t which_decomp, double* truncerr,
                    double* eigs, int64_t cap, int64_t* neigs) {
  return guard(e->env.ctx, [&] {
    FactorizeParams prm;
    prm.ortho_left = ortho_left;
    prm.maxdim = maxdim <= 0 ? INT64_MAX : maxdim;
    prm.mindim = mindim;
    prm.cutoff = cutoff;
    prm.noise = noise;
    prm.which = which_decomp & 15;
    prm.svd_alg = which_decomp >> 4;
    TensorP vq = phi->t->nrow == 1 ? phi->t : relayout(e->env.ctx, *phi->t, 1);
    FactorizeResult f = e->env.replacebond(pos, *vq, prm, normalize != 0);
    *truncerr = f.truncerr;
    *neigs = (int64_t)f.eigs.size();
    for (int64_t i = 0; i < std::min<int64_t>(cap, *neigs); i++) eigs[i] = f.eigs[i];
  });
}

int tnl_svd_split(tnl_env_t e, int32_t pos, tnl_tensor_t phi, int32_t ortho_left, int64_t maxdim, int64_t mindim,
                  double cutoff, int32_t normalize, int32_t svd_alg, double* truncerr, double* eigs, int64_t cap,
                  int64_t* neigs, tnl_tensor_t* carry) {
  return guard(e->env.ctx, [&] {
    FactorizeParams prm;
    prm.ortho_left = ortho_left;
    prm.maxdim = maxdim <= 0 ? INT64_MAX : maxdim;
    prm.mindim = mindim;
    prm.cutoff = cutoff;
    prm.svd_alg = svd_alg;
    FactorizeResult f = e->env.svd_split(pos, *phi->t, prm, normalize != 0, carry == nullptr);
    if (carry) *carry = new tnl_tensor_s{ortho_left ? f.R : f.L};
    *truncerr = f.truncerr;
    *neigs = (int64_t)f.eigs.size();
    for (int64_t i = 0; i < std::min<int64_t>(cap, *neigs); i++) eigs[i] = f.eigs[i];
  });
}


/* ---- generic tensor algebra (algebra.hpp) ---------------------------------------------------------------------- */
int tnl_tensor_permute(tnl_tensor_t t, const int32_t* perm, int32_t nrow, tnl_tensor_t* out) {
  return guard(t->t->ctx, [&] {
    std::vector<int> p(perm, perm + t->t->rank());
    *out = new tnl_tensor_s{permute(t->t->ctx, *t->t, p, nrow)};
  });
}
int tnl_tensor_dag(tnl_tensor_t t, tnl_tensor_t* out) {
  return guard(t->t->ctx, [&] { *out = new tnl_tensor_s{dag_copy(t->t->ctx, *t->t)}; });
}
int tnl_tensor_contract(tnl_tensor_t a, const int32_t* labels_a, int32_t dag_a, tnl_tensor_t b, const int32_t* labels_b,
                        int32_t dag_b, tnl_tensor_t* out, int32_t* labels_out, int32_t* rank_out) {
  return guard(a->t->ctx, [&] {
    std::vector<int> la(labels_a, labels_a + a->t->rank()), lb(labels_b, labels_b + b->t->rank()), lc;
    TensorP c = contract(a->t->ctx, *a->t, la, dag_a != 0, *b->t, lb, dag_b != 0, &lc);
    a->t->ctx->sync();
    for (size_t k = 0; k < lc.size(); k++) labels_out[k] = lc[k];
    *rank_out = (int32_t)lc.size();
    *out = new tnl_tensor_s{c};
  });
}
int tnl_tensor_directsum(tnl_tensor_t a, int32_t ia, tnl_tensor_t b, int32_t ib, tnl_tensor_t* out) {
  return guard(a->t->ctx, [&] {
    *out = new tnl_tensor_s{directsum(a->t->ctx, *a->t, ia, *b->t, ib)};
    a->t->ctx->sync();
  });
}
int tnl_tensor_factorize(tnl_tensor_t t, int32_t nleft, int32_t ortho_left, int64_t maxdim, int64_t mindim, double cutoff,
                         int32_t which_decomp, tnl_tensor_t* L, tnl_tensor_t* R, double* truncerr, double* eigs, int64_t cap,
                         int64_t* neigs) {
  return guard(t->t->ctx, [&] {
    Ctx* ctx = t->t->ctx;
    TNL_CHECK(nleft >= 1 && nleft < t->t->rank(), "factorize needs a proper bipartition");
    FactorizeParams prm;
    prm.ortho_left = ortho_left;
    prm.maxdim = maxdim <= 0 ? INT64_MAX : maxdim;
    prm.mindim = mindim;
    prm.cutoff = cutoff;
    prm.which = which_decomp & 15;
    prm.svd_alg = which_decomp >> 4;
    TensorP T = t->t->nrow == nleft ? t->t : relayout(ctx, *t->t, nleft);
    FactorizeResult f = factorize(ctx, *T, prm);
    ctx->sync();
    *L = new tnl_tensor_s{f.L};
    *R = new tnl_tensor_s{f.R};
    if (truncerr) *truncerr = f.truncerr;
    if (neigs) *neigs = (int64_t)f.eigs.size();
    if (eigs) for (int64_t i = 0; i < std::min<int64_t>(cap, (int64_t)f.eigs.size()); i++) eigs[i] = f.eigs[i];
  });
}

int tnl_sumop_create(tnl_ctx_t c, int32_t rank, const int32_t* vlabels, tnl_sumop_t* out) {
  return guard(&c->ctx, [&] {
    auto* h = new tnl_sumop_s(&c->ctx);
    h->op.vlabels.assign(vlabels, vlabels + rank);
    *out = h;
  });
}
int tnl_sumop_destroy(tnl_sumop_t op) {
  return guard(nullptr, [&] { delete op; });
}
int tnl_sumop_add_term(tnl_sumop_t op, int32_t nops, const tnl_tensor_t* tensors, const int32_t* labels_flat) {
  return guard(op->op.ctx, [&] {
    SumOp::Term term;
    int at = 0;
    for (int k = 0; k < nops; k++) {
      SumOp::Operand o;
      o.t = tensors[k]->t;
      o.labels.assign(labels_flat + at, labels_flat + at + o.t->rank());
      at += o.t->rank();
      term.ops.push_back(std::move(o));
    }
    op->op.terms.push_back(std::move(term));
  });
}
int tnl_sumop_set_relabel(tnl_sumop_t op, int32_t n, const int32_t* from, const int32_t* to) {
  return guard(op->op.ctx, [&] {
    op->op.relabel.clear();
    for (int k = 0; k < n; k++) op->op.relabel.emplace_back(from[k], to[k]);
  });
}
int tnl_sumop_add_projector(tnl_sumop_t op, tnl_tensor_t m, double weight) {
  return guard(op->op.ctx, [&] {
    TNL_CHECK(weight > 0.0, "`weight` parameter should be > 0.0");
    op->op.projs.push_back(m->t);
    op->op.weight = weight;
  });
}
int tnl_sumop_apply(tnl_sumop_t op, tnl_tensor_t v, tnl_tensor_t* out) {
  return guard(op->op.ctx, [&] {
    auto o = std::make_shared<Tensor>(op->op.ctx, v->t->inds, v->t->nrow, true, v->t->cplx);
    op->op.apply(*v->t, *o);
    op->op.ctx->sync();
    *out = new tnl_tensor_s{o};
  });
}
int tnl_sumop_apply_flops(tnl_sumop_t op, double* flops) {
  return guard(op->op.ctx, [&] { *flops = op->op.flops; });
}
int tnl_sumop_eigsolve(tnl_sumop_t op, tnl_tensor_t phi, double tol, int32_t krylovdim, int32_t maxiter, int32_t eager,
                       double* eval, int32_t* converged, int32_t* numops, int32_t* numiter, double* normres) {
  return guard(op->op.ctx, [&] {
    LanczosResult r = krylov_eigsolve(op->op.ctx, op->op, *phi->t, tol, krylovdim, maxiter, eager != 0);
    *eval = r.eval; *converged = r.converged; *numops = r.numops; *numiter = r.numiter; *normres = r.normres;
  });
}
int tnl_sumop_exponentiate(tnl_sumop_t op, tnl_tensor_t phi, double t_re, double t_im, double tol, int32_t krylovdim,
                           int32_t maxiter, int32_t eager, int32_t* converged, int32_t* numops, int32_t* numiter, double* err) {
  return guard(op->op.ctx, [&] {
    if (t_im != 0.0) phi->t = promoted(phi->t);
    ExpResult r = krylov_exponentiate(op->op.ctx, op->op, *phi->t, t_re, t_im, tol, krylovdim, maxiter, eager != 0);
    *converged = r.converged; *numops = r.numops; *numiter = r.numiter; *err = r.err;
  });
}

}  // extern "C"
