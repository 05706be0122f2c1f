// On-device bond factorisation with ITensors truncation semantics.
//
// Replaces ITensorMPS `replacebond!` -> ITensors `factorize` -> NDTensors block-sparse `svd` / `eigen`
// (third-party; reached from src/mps/update_site.jl:64-76 and :162-168).  In the charge-fused layout
// with nrow = split every charge group of the tensor already IS the dense matrix NDTensors would
// decompose, so the per-block LAPACK calls become one cuSOLVER call per charge group:
//   * "svd"   path (reference: cutoff <= 1e-12 and no noise):  cusolverDnDgesvd (QR iteration)
//   * "eigen" path (reference: noise > 0 or cutoff > 1e-12):   rho = M M^T (+ noise * X X^T) with the
//             grouped DGEMM kernel, cusolverDnDsyevd, other factor by projection (factorize_eigen).
// Truncation: pooled spectrum sorted descending -> NDTensors `truncate!` (relative cumulative cutoff,
// maxdim/mindim, degeneracy-aware docut) -> per-group keep `value > docut`; groups keeping nothing drop.
#include <cuda.h>
#include <cusolverDn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <thread>

#include "env.hpp"

namespace tnl {

#define CUSOLVER_OK(call)                                                                              \
  do {                                                                                                 \
    cusolverStatus_t s__ = (call);                                                                     \
    if (s__ != CUSOLVER_STATUS_SUCCESS)                                                                \
      throw ::tnl::Error(4, std::string("cuSOLVER error ") + std::to_string((int)s__) + " in " #call); \
  } while (0)

static cusolverDnHandle_t solver(Ctx* ctx) {
  if (!ctx->cusolver) {
    cusolverDnHandle_t h;
    CUSOLVER_OK(cusolverDnCreate(&h));
    CUSOLVER_OK(cusolverDnSetStream(h, ctx->stream));
    ctx->cusolver = h;
  }
  return (cusolverDnHandle_t)ctx->cusolver;
}
// ---------------------------------------------------------------------------------------------------
// SM partitions for the per-charge-group eigendecompositions.
// cusolverDnDsyevd is a latency chain -- its tridiagonalisation costs ~11 us per COLUMN whatever the machine size,
// plus a bandwidth term ~ n^3 / #SMs (measured on B200, profiles/r02g_eigh_green.json:  t[ms] ~ 0.0112 n + 9.2e-8 n^3 / S)
// -- and its persistent kernel sizes itself to the SMs it sees, so decompositions on plain side streams do not
// overlap (profiles/r02c_eigh_bench.json).  Inside GREEN CONTEXTS (disjoint SM partitions, CUDA driver API) they do:
// the 11 charge groups of the chi = 4096 bench bond take 196 ms one after the other and ~100 ms on three partitions.
// A partition set is a tuple of sizes in units of 16 SMs (on B200 the driver hands out nine 16-SM groups + 4 SMs; with 8-SM
// groups it only places fifteen, profiles/r02g_eigh_green.json); the batch driver
// below picks, per call, the set with the smallest modelled makespan out of a short candidate list.
struct PartSolver {
  cudaStream_t s = nullptr; cusolverDnHandle_t h = nullptr; double* work = nullptr; size_t bytes = 0; int* info = nullptr;
  int sms = 0;
};
constexpr int kPartUnitSMs = 16;
struct PartSet { std::vector<int> units; std::vector<PartSolver> parts; bool ok = false; };

struct GreenApi {
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  bool ok = false;
};
static GreenApi& green_api() {
  static GreenApi api = [] {
    GreenApi a;
    auto get = [](const char* name) -> void* {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
      return p;
    };
    a.DeviceGet = (decltype(a.DeviceGet))get("cuDeviceGet");
    a.DeviceGetDevResource = (decltype(a.DeviceGetDevResource))get("cuDeviceGetDevResource");
    a.DevSmResourceSplitByCount = (decltype(a.DevSmResourceSplitByCount))get("cuDevSmResourceSplitByCount");
    a.DevResourceGenerateDesc = (decltype(a.DevResourceGenerateDesc))get("cuDevResourceGenerateDesc");
    a.GreenCtxCreate = (decltype(a.GreenCtxCreate))get("cuGreenCtxCreate");
    a.GreenCtxStreamCreate = (decltype(a.GreenCtxStreamCreate))get("cuGreenCtxStreamCreate");
    a.ok = a.DeviceGet && a.DeviceGetDevResource && a.DevSmResourceSplitByCount && a.DevResourceGenerateDesc && a.GreenCtxCreate &&
           a.GreenCtxStreamCreate;
    cudaGetLastError();
    return a;
  }();
  return api;
}

// partition set `units` (sizes in units of 16 SMs) on `device`; one unit tuple {0} = the whole device on a plain stream.
// Sets are created once per (device, tuple) and shared by every context on that device.
static PartSet& part_set(int device, const std::vector<int>& units) {
  static std::map<std::pair<int, std::vector<int>>, PartSet> all;
  auto key = std::make_pair(device, units);
  auto it = all.find(key);
  if (it != all.end()) return it->second;
  PartSet& ps = all[key];
  ps.units = units;
  auto finish = [&](PartSolver& sv) {
    CUSOLVER_OK(cusolverDnCreate(&sv.h));
    CUSOLVER_OK(cusolverDnSetStream(sv.h, sv.s));
    CUDA_OK(cudaMalloc(&sv.info, 64 * sizeof(int)));
    // first call on a handle pays its lazy initialisation (internal buffers, kernel loading): do it here, on a 96 x 96
    // identity, instead of inside the first real decomposition
    const int n = 96;
    double *A = nullptr, *Wv = nullptr, *wk = nullptr;
    int lw = 0;
    CUDA_OK(cudaMalloc(&A, sizeof(double) * n * n));
    CUDA_OK(cudaMalloc(&Wv, sizeof(double) * n));
    std::vector<double> eye((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) eye[(size_t)i * n + i] = 1.0 + i;
    CUDA_OK(cudaMemcpy(A, eye.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CUSOLVER_OK(cusolverDnDsyevd_bufferSize(sv.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, Wv, &lw));
    CUDA_OK(cudaMalloc(&wk, sizeof(double) * std::max(lw, 1)));
    CUSOLVER_OK(cusolverDnDsyevd(sv.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, Wv, wk, lw, sv.info));
    CUDA_OK(cudaStreamSynchronize(sv.s));
    cudaFree(A); cudaFree(Wv); cudaFree(wk);
  };
  if (units.size() == 1) {                         // whole device
    ps.parts.resize(1);
    CUDA_OK(cudaStreamCreateWithFlags(&ps.parts[0].s, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    ps.parts[0].sms = prop.multiProcessorCount;
    finish(ps.parts[0]);
    ps.ok = true;
    return ps;
  }
  GreenApi& api = green_api();
  static const bool dbg = getenv("TNL_EIGH_DEBUG") != nullptr;
  auto fail = [&](const char* what, int rc) -> PartSet& {
    if (dbg) fprintf(stderr, "[tnl] SM partition set unavailable: %s (rc %d)\n", what, rc);
    return ps;
  };
  if (!api.ok) return fail("green-context driver entry points missing", 0);
  CUdevice dev;
  CUdevResource all_sm, rem;
  CUresult rc;
  if ((rc = api.DeviceGet(&dev, device)) != CUDA_SUCCESS) return fail("cuDeviceGet", rc);
  if ((rc = api.DeviceGetDevResource(dev, &all_sm, CU_DEV_RESOURCE_TYPE_SM)) != CUDA_SUCCESS) return fail("cuDeviceGetDevResource", rc);
  int total_units = 0;
  for (int u : units) total_units += u;
  std::vector<CUdevResource> res((size_t)total_units);
  unsigned int ng = (unsigned int)total_units;
  rc = api.DevSmResourceSplitByCount(res.data(), &ng, &all_sm, &rem, 0, kPartUnitSMs);
  if (rc != CUDA_SUCCESS || (int)ng < total_units) return fail("cuDevSmResourceSplitByCount (16-SM units)", rc != CUDA_SUCCESS ? (int)rc : -(int)ng);
  ps.parts.resize(units.size());
  int at = 0;
  for (size_t p = 0; p < units.size(); p++) {
    CUdevResourceDesc desc;
    CUgreenCtx g;
    CUstream st;
    if ((rc = api.DevResourceGenerateDesc(&desc, &res[at], (unsigned int)units[p])) != CUDA_SUCCESS) return fail("cuDevResourceGenerateDesc", rc);
    if ((rc = api.GreenCtxCreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) return fail("cuGreenCtxCreate", rc);
    if ((rc = api.GreenCtxStreamCreate(&st, g, CU_STREAM_NON_BLOCKING, 0)) != CUDA_SUCCESS) return fail("cuGreenCtxStreamCreate", rc);
    ps.parts[p].s = (cudaStream_t)st;
    ps.parts[p].sms = 0;
    for (int u = 0; u < units[p]; u++) ps.parts[p].sms += (int)res[at + u].sm.smCount;
    at += units[p];
    finish(ps.parts[p]);
  }
  ps.ok = true;
  return ps;
}
// modelled milliseconds of one cusolverDn{D,Z}syevd / heevd of order n on S SMs (fit of profiles/r02g_eigh_green.json)
static double eigh_cost_ms(double n, double S, bool cplx) { return (0.0112 * n + 9.2e-8 * n * n * n / S) * (cplx ? 2.0 : 1.0); }

static double* solver_ws(Ctx* ctx, size_t doubles) {
  size_t bytes = doubles * sizeof(double);
  if (bytes > ctx->solver_work_bytes) {
    ctx->sync();
    if (ctx->solver_work) cudaFree(ctx->solver_work);
    CUDA_OK(cudaMalloc(&ctx->solver_work, bytes));
    ctx->solver_work_bytes = bytes;
  }
  return (double*)ctx->solver_work;
}

// NDTensors `truncate!` on a descending spectrum; returns number kept, truncerr, docut
static int64_t truncate_spectrum(std::vector<double>& P, int64_t maxdim, int64_t mindim, double cutoff,
                                 double& truncerr, double& docut) {
  const int64_t origm = (int64_t)P.size();
  truncerr = 0.0;
  docut = 0.0;
  if (origm == 0) return 0;
  if (origm == 1) { docut = std::fabs(P[0]) / 2; return 1; }
  maxdim = std::min(maxdim, origm);
  double s = P[0] < 0 ? -1.0 : (P[0] > 0 ? 1.0 : 0.0);
  if (s < 0) for (auto& x : P) x *= s;
  for (int64_t n = origm - 1; n >= 0; n--) {
    if (P[n] >= 0) break;
    P[n] = 0.0;
  }
  int64_t n = origm;
  while (n > maxdim) { truncerr += P[n - 1]; n--; }
  double scale = 0.0;
  for (double x : P) scale += x;
  if (scale == 0.0) scale = 1.0;
  while (n > mindim && (truncerr + P[n - 1] <= cutoff * scale)) { truncerr += P[n - 1]; n--; }
  truncerr /= scale;
  if (n < 1) n = 1;
  if (n < origm) {
    docut = (P[n - 1] + P[n]) / 2;
    if (std::fabs(P[n - 1] - P[n]) < 1e-3 * P[n - 1]) docut += 1e-3 * P[n - 1];
  }
  if (s < 0) for (auto& x : P) x *= s;
  P.resize(n);
  return n;
}

__global__ void gather_cols_kernel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                                   int64_t R, int nk, const int* __restrict__ idx, const double* __restrict__ scale) {
  int64_t n = R * nk;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e % R, c = e / R;
    double v = src[(int64_t)idx[c] * lds + r];
    if (scale) v *= scale[c];
    dst[c * ldd + r] = v;
  }
}
// dst[R x nk] = src[:, idx[c]] (* scale[c])
static void gather_cols(Ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t R,
                        const std::vector<int>& idx, const std::vector<double>* scale) {
  int nk = (int)idx.size();
  if (R * nk == 0) return;
  int* d_idx = ctx->upload(idx);
  double* d_sc = scale ? ctx->upload(*scale) : nullptr;
  int grid = (int)std::min<int64_t>((R * nk + 255) / 256, 1184);
  gather_cols_kernel<<<grid, 256, 0, ctx->stream>>>(dst, ldd, src, lds, R, nk, d_idx, d_sc);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->free(d_idx);
  ctx->free(d_sc);
}

__global__ void scatter_cols_kernel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                                    int64_t R, int nk, const int* __restrict__ idx) {
  int64_t n = R * nk;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e % R, c = e / R;
    dst[(int64_t)idx[c] * ldd + r] = src[c * lds + r];
  }
}
// dst[:, idx[c]] = src[:, c]
static void scatter_cols(Ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t R,
                         const std::vector<int>& idx) {
  int nk = (int)idx.size();
  if (R * nk == 0) return;
  int* d_idx = ctx->upload(idx);
  int grid = (int)std::min<int64_t>((R * nk + 255) / 256, 1184);
  scatter_cols_kernel<<<grid, 256, 0, ctx->stream>>>(dst, ldd, src, lds, R, nk, d_idx);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->free(d_idx);
}

struct FG {                     // one charge group of the factorisation
  Charge q;
  int64_t R = 0, C = 0;
  const double* M = nullptr; int64_t ldm = 0;     // data of T (nullptr: group opened by the noise term only)
  struct NoiseOp { const double* X = nullptr; int64_t ldx = 0, XK = 0; int64_t xn = 0; /* plane distance (complex X) */ };
  std::vector<NoiseOp> Xs;                        // noise operands, one slot per MPO term (X == nullptr: none)
  // results
  double* U = nullptr; int64_t ldu = 0;           // left vectors  [R x k]
  double* Vt = nullptr; int64_t ldv = 0;          // right vectors [k x C]  (svd path)
  double* E = nullptr;                            // eigenvectors (eigen path) [n x n]
  std::vector<double> vals;                       // sigma (svd) or eigenvalues (eigen), LAPACK order
  std::vector<int> keep;                          // kept columns, descending weight
  bool mine = true;                               // multi-GPU: this rank decomposes the group
  int64_t nvals = 0;                              // number of spectrum values the group contributes
  int64_t kmax = INT64_MAX;                       // Gram driver: only the min(R, C) largest eigenvalues of M M^T are
                                                  // singular values, the rest is structurally zero and never kept
  // weights (|value|) of the group's admissible spectrum entries, descending, with their positions in `vals`
  std::vector<int> admissible() const {
    std::vector<int> order(vals.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::fabs(vals[a]) > std::fabs(vals[b]); });
    if ((int64_t)order.size() > kmax) order.resize((size_t)kmax);
    return order;
  }
};

static const std::vector<std::vector<int>> kPartCandidates = {{0}, {8, 1}, {7, 2}, {5, 4}, {4, 3, 2}, {3, 3, 3}, {4, 2, 2, 1}, {2, 2, 2, 2, 1}};
// modelled makespan of a list of decompositions (orders n) on the partition set `units`: largest first, each job to
// the partition on which it would finish earliest; `where` (optional) receives the partition of every job
static double eigh_makespan(const std::vector<int64_t>& n, const std::vector<char>& cplx, const std::vector<int>& units, int device_sms,
                            std::vector<int>* where) {
  std::vector<int> sms;
  if (units.size() == 1) sms = {device_sms}; else for (int x : units) sms.push_back(kPartUnitSMs * x);
  std::vector<size_t> order(n.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return n[a] > n[b]; });
  std::vector<double> fin(sms.size(), 0.0);
  if (where) where->assign(n.size(), 0);
  for (size_t k : order) {
    size_t best = 0;
    double bt = 1e300;
    for (size_t p = 0; p < sms.size(); p++) {
      const double t = fin[p] + eigh_cost_ms((double)n[k], sms[p], cplx.empty() ? false : cplx[k] != 0);
      if (t < bt) { bt = t; best = p; }
    }
    fin[best] = bt;
    if (where) (*where)[k] = (int)best;
  }
  return fin.empty() ? 0.0 : *std::max_element(fin.begin(), fin.end());
}
// smallest modelled makespan over the candidate partition sets (what syevd_batch will achieve for these jobs)
static double eigh_best_makespan(const std::vector<int64_t>& n, int device_sms) {
  double best = 1e300;
  for (auto& u : kPartCandidates) best = std::min(best, eigh_makespan(n, {}, u, device_sms, nullptr));
  return n.empty() ? 0.0 : best;
}

// Hermitian eigendecompositions of independent dense matrices (one per charge group), spread over SM partitions
// (green contexts, see part_set) and host threads: cusolverDnDsyevd is host-driven (its calls block while panels are
// factorised), so each partition gets its own host thread.  Jobs go, largest first, to the partition on which they
// would finish earliest under the cost model; the partition set is the candidate with the smallest modelled makespan.
struct EighJob { int64_t n; double* A; double* W; std::vector<double>* vals; int64_t lda; bool cplx = false; };   // A: in matrix / out eigenvectors (cplx: interleaved)
static void syevd_batch(Ctx* ctx, std::vector<EighJob>& jobs) {
  if (jobs.empty()) return;
  ctx->sync();
  std::vector<size_t> order(jobs.size());
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return jobs[a].n > jobs[b].n; });
  // ---- choose the partition set
  std::vector<std::vector<int>> cands;
  if (const char* e = getenv("TNL_EIGH_PARTS")) {            // "0" = whole device only, "4,2,2,1" = this set, unset = auto
    std::vector<int> u;
    for (const char* q = e; *q;) { u.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q == ',') q++; }
    cands.push_back(u.empty() ? std::vector<int>{0} : u);
  } else {
    cands = kPartCandidates;
  }
  std::vector<int64_t> jn;
  std::vector<char> jc;
  for (auto& j : jobs) { jn.push_back(j.n); jc.push_back(j.cplx ? 1 : 0); }
  {
    // create every candidate set (green contexts, streams, cuSOLVER handles) the first time a batch with a large
    // decomposition arrives: the set changes from bond to bond, and a lazily created one would put its one-off cost
    // (hundreds of milliseconds) into the middle of a sweep
    static std::map<int, bool> warmed;
    if (!warmed[ctx->device] && jobs[order[0]].n >= 512 && !getenv("TNL_EIGH_PARTS")) {
      warmed[ctx->device] = true;
      for (auto& u : kPartCandidates) part_set(ctx->device, u);
    }
  }
  PartSet* ps = nullptr;
  std::vector<int> where;
  {
    double best = 1e300;
    for (auto& u : cands) {
      std::vector<int> w;
      const double t = eigh_makespan(jn, jc, u, ctx->num_sms, &w);
      if (t < best * 0.97) {                       // a partitioned set must win by a margin: it costs extra launches
        PartSet& cand = part_set(ctx->device, u.size() == 1 ? std::vector<int>{0} : u);
        if (!cand.ok) continue;
        best = t; ps = &cand; where = w;
      }
    }
    if (!ps) {
      ps = &part_set(ctx->device, {0});
      eigh_makespan(jn, jc, {0}, ctx->num_sms, &where);
    }
  }
  if (getenv("TNL_EIGH_DEBUG")) {
    fprintf(stderr, "[tnl] syevd_batch: %zu jobs (largest n = %lld) on partition set {", jobs.size(), (long long)jobs[order[0]].n);
    for (auto& sv : ps->parts) fprintf(stderr, " %d", sv.sms);
    fprintf(stderr, " } SMs\n");
  }
  auto& sides = ps->parts;
  std::vector<int> slot_of(jobs.size());
  std::vector<int> used(sides.size(), 0);
  for (size_t k : order) {
    TNL_CHECK(used[where[k]] < 64, "too many charge groups per partition");
    slot_of[k] = used[where[k]]++;
  }
  std::vector<std::string> errs(sides.size());
  const auto t_batch0 = std::chrono::steady_clock::now();
  auto worker = [&](size_t si) {
    try {
      CUDA_OK(cudaSetDevice(ctx->device));
      PartSolver& sv = sides[si];
      for (size_t k : order) {
        if ((size_t)where[k] != si) continue;
        EighJob& j = jobs[k];
        int lwork = 0;
        if (j.cplx)
          CUSOLVER_OK(cusolverDnZheevd_bufferSize(sv.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)j.n, (cuDoubleComplex*)j.A,
                                                  (int)j.lda, j.W, &lwork));
        else
          CUSOLVER_OK(cusolverDnDsyevd_bufferSize(sv.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)j.n, j.A, (int)j.lda,
                                                  j.W, &lwork));
        const size_t need = (size_t)lwork * sizeof(double) * (j.cplx ? 2 : 1);
        if (need > sv.bytes) {
          CUDA_OK(cudaStreamSynchronize(sv.s));
          if (sv.work) cudaFree(sv.work);
          sv.bytes = need * 5 / 4;
          CUDA_OK(cudaMalloc(&sv.work, sv.bytes));
        }
        if (j.cplx)
          CUSOLVER_OK(cusolverDnZheevd(sv.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)j.n, (cuDoubleComplex*)j.A, (int)j.lda,
                                       j.W, (cuDoubleComplex*)sv.work, lwork, sv.info + slot_of[k]));
        else
          CUSOLVER_OK(cusolverDnDsyevd(sv.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)j.n, j.A, (int)j.lda, j.W,
                                       sv.work, lwork, sv.info + slot_of[k]));
      }
      CUDA_OK(cudaStreamSynchronize(sv.s));
    } catch (const std::exception& e) {
      errs[si] = e.what();
    }
  };
  {
    std::vector<std::thread> th;
    for (size_t si = 1; si < sides.size(); si++)
      if (used[si] > 0) th.emplace_back(worker, si);
    worker(0);
    for (auto& t : th) t.join();
    for (auto& e : errs) TNL_CHECK(e.empty(), e);
  }
  for (auto& sv : sides) CUDA_OK(cudaStreamSynchronize(sv.s));
  if (getenv("TNL_EIGH_DEBUG"))
    fprintf(stderr, "[tnl] syevd_batch: %.2f ms wall\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_batch0).count());
  for (size_t k : order) {
    EighJob& j = jobs[k];
    PartSolver& sv = sides[where[k]];
    j.vals->resize(j.n);
    CUDA_OK(cudaMemcpy(j.vals->data(), j.W, j.n * sizeof(double), cudaMemcpyDeviceToHost));
    int info = 0;
    CUDA_OK(cudaMemcpy(&info, sv.info + slot_of[k], sizeof(int), cudaMemcpyDeviceToHost));
    TNL_CHECK(info == 0, "cuSOLVER syevd / heevd did not converge");
  }
}

// ---------------------------------------------------------------------------------------------
// QR / LQ gauge move without truncation (ITensors `factorize(...; which_decomp = "qr")`, what
// `orthogonalize!` needs): ortho left  M = Q R  -> L = Q [R x k], R = R [k x C];
//                          ortho right M = L Q  -> via QR of M^T.                k = min(R, C).
__global__ void triu_kernel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                            int64_t k, int64_t n, bool transpose_out) {
  // dst = triu(src[0:k, 0:n])  (or its transpose [n x k])
  int64_t tot = k * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = e % k, j = e / k;
    double v = (i <= j) ? src[j * lds + i] : 0.0;
    if (transpose_out) dst[i * ldd + j] = v; else dst[j * ldd + i] = v;
  }
}

static FactorizeResult factorize_qr(Ctx* ctx, const Tensor& T, bool left, int new_dir_on_L) {
  cusolverDnHandle_t H = solver(ctx);
  const int split = T.nrow;
  Index m;
  m.nq = T.inds[0].nq;
  m.dir = new_dir_on_L;
  for (const Group& g : T.groups) {
    int64_t k = std::min(g.R, g.C);
    if (k == 0) continue;
    m.dims.push_back((int)k);
    m.qns.push_back(g.q);
  }
  std::vector<Index> li(T.inds.begin(), T.inds.begin() + split), ri;
  li.push_back(m);
  ri.push_back(m.dag());
  ri.insert(ri.end(), T.inds.begin() + split, T.inds.end());
  FactorizeResult res;
  res.path = "qr";
  res.L = std::make_shared<Tensor>(ctx, li, split);
  res.R = std::make_shared<Tensor>(ctx, ri, 1);
  std::vector<void*> temps;
  for (const Group& g : T.groups) {
    const int64_t R = g.R, C = g.C, k = std::min(R, C);
    if (k == 0) continue;
    const Group& GL = res.L->groups[res.L->find_group(g.q)];
    const Group& GR = res.R->groups[res.R->find_group(g.q)];
    const int64_t mm = left ? R : C, nn = left ? C : R;          // QR of the [mm x nn] matrix
    double* Aw = (double*)ctx->alloc(mm * nn * sizeof(double));
    temps.push_back(Aw);
    if (left) copy2d(ctx, Aw, mm, T.d + g.base, g.ld, R, C);
    else transpose(ctx, Aw, mm, T.d + g.base, g.ld, R, C);       // M^T
    double* tau = (double*)ctx->alloc(k * sizeof(double));
    temps.push_back(tau);
    int lw1 = 0, lw2 = 0;
    CUSOLVER_OK(cusolverDnDgeqrf_bufferSize(H, (int)mm, (int)nn, Aw, (int)mm, &lw1));
    CUSOLVER_OK(cusolverDnDorgqr_bufferSize(H, (int)mm, (int)k, (int)k, Aw, (int)mm, tau, &lw2));
    double* work = solver_ws(ctx, (size_t)std::max(lw1, lw2));
    CUSOLVER_OK(cusolverDnDgeqrf(H, (int)mm, (int)nn, Aw, (int)mm, tau, work, lw1, ctx->d_info));
    // triangular factor [k x nn]
    int grid = (int)std::min<int64_t>((k * nn + 255) / 256, 1184);
    if (left) triu_kernel<<<grid, 256, 0, ctx->stream>>>(res.R->d + GR.base, GR.ld, Aw, mm, k, nn, false);   // R [k x C]
    else triu_kernel<<<grid, 256, 0, ctx->stream>>>(res.L->d + GL.base, GL.ld, Aw, mm, k, nn, true);          // (R')^T [R x k]
    CUDA_OK(cudaGetLastError());
    ctx->cnt.launches++;
    CUSOLVER_OK(cusolverDnDorgqr(H, (int)mm, (int)k, (int)k, Aw, (int)mm, tau, work, lw2, ctx->d_info));
    if (left) copy2d(ctx, res.L->d + GL.base, GL.ld, Aw, mm, R, k);                                           // Q [R x k]
    else transpose(ctx, res.R->d + GR.base, GR.ld, Aw, mm, C, k);                                              // Q^T [k x C]
    int info = 0;
    CUDA_OK(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    TNL_CHECK(info == 0, "cuSOLVER QR failed");
  }
  ctx->sync();
  for (void* p : temps) ctx->free(p);
  return res;
}

// ---------------------------------------------------------------------------------------------------
// ComplexF64 (planar) tensors: Hermitian-eigenproblem route only.  rho = M M^+ (ortho left) or M^+ M (right)
// with four launches of the real grouped DGEMM, planar -> interleaved, cusolverDnZheevd per charge group,
// eigenvectors back to planes, the other factor by projection.  Covers `factorize`'s eigen path and -- as in the
// "gram" driver of the real code -- its SVD path (same kept spectrum sigma^2, same truncation); the dedicated
// complex SVD / QR drivers and the noise term are not built.
__global__ void interleave_kernel(double2* __restrict__ z, const double* __restrict__ re, const double* __restrict__ im, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    z[e] = make_double2(re[e], im[e]);
}
__global__ void deinterleave_kernel(double* __restrict__ re, double* __restrict__ im, const double2* __restrict__ z, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = z[e];
    re[e] = v.x;
    im[e] = v.y;
  }
}

static FactorizeResult factorize_complex(Ctx* ctx, const Tensor& T, const FactorizeParams& prm) {
  const int split = T.nrow;
  const bool noisy = prm.noise != 0.0 && prm.noiseX != nullptr && prm.which != 3;
  // which == 3 (the untruncated gauge move of `qr` / `orthogonalize!`): the isometry spans range(M) -- the leading
  // min(R, C) eigenvectors of rho per charge group -- and the other factor is its projection; another gauge of the
  // same factorisation a Householder QR would give
  const bool gauge_only = prm.which == 3;
  const bool left = prm.ortho_left != 0;
  cusolverDnHandle_t H = solver(ctx);
  const int64_t tn = T.nelem;
  std::vector<FG> fg;
  for (const Group& g : T.groups) {
    FG f; f.q = g.q; f.R = g.R; f.C = g.C; f.M = T.d + g.base; f.ldm = g.ld;
    fg.push_back(f);
  }
  // noise operands (noiseterm, update_site.jl:59-62): rho += noise * sum_t X_t X_t^+ (ortho left) / X_t^+ X_t (right);
  // a sector the operands open on the kept side only exists on the factor that carries the eigenvectors
  std::vector<const Tensor*> nops;
  if (noisy) {
    nops.push_back(prm.noiseX);
    for (const Tensor* x : prm.noiseXmore) nops.push_back(x);
    for (size_t t = 0; t < nops.size(); t++) {
      const Tensor& X = *nops[t];
      for (const Group& g : X.groups) {
        auto it = std::find_if(fg.begin(), fg.end(), [&](const FG& f) { return f.q == g.q; });
        if (it == fg.end()) {
          FG f; f.q = g.q;
          if (left) { f.R = g.R; f.C = 0; } else { f.R = 0; f.C = g.C; }
          fg.push_back(f);
          it = fg.end() - 1;
        }
        if (left) TNL_CHECK(it->R == g.R, "noise operand rows do not match");
        else TNL_CHECK(it->C == g.C, "noise operand cols do not match");
        it->Xs.resize(nops.size());
        it->Xs[t].X = X.d + g.base; it->Xs[t].ldx = g.ld; it->Xs[t].XK = left ? g.C : g.R;
        it->Xs[t].xn = X.cplx ? X.nelem : 0;
      }
    }
    std::sort(fg.begin(), fg.end(), [](const FG& a, const FG& b) { return a.q < b.q; });
  }
  std::vector<void*> temps;
  auto talloc = [&](int64_t n) {
    n = (std::max<int64_t>(n, 1) + 1) & ~int64_t(1);
    double* p = (double*)ctx->alloc((size_t)n * sizeof(double));
    CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(double), ctx->stream));
    temps.push_back(p);
    return p;
  };
  auto off = [](const double* p) { return (int64_t)(reinterpret_cast<intptr_t>(p) / (intptr_t)sizeof(double)); };
  // ---- rho per group (planes Er, Ei with leading dimension n)
  std::vector<double*> Er(fg.size(), nullptr), Ei(fg.size(), nullptr);
  {
    std::vector<GemmProblem> pr;
    std::vector<size_t> which_g;
    for (size_t gi = 0; gi < fg.size(); gi++) {
      FG& f = fg[gi];
      const int64_t n = left ? f.R : f.C;
      if (n == 0) continue;
      Er[gi] = talloc(n * n);
      Ei[gi] = talloc(n * n);
      if (!f.M) continue;                        // opened by the noise term only: rho starts at zero
      GemmProblem p{};
      p.M = p.N = (int)n; p.ldc = (int)n; p.K = (int)(left ? f.C : f.R);
      p.a = p.b = off(f.M); p.lda = p.ldb = (int)f.ldm;
      p.c = off(Er[gi]);
      pr.push_back(p);
      which_g.push_back(gi);
    }
    // the real and imaginary result planes live in separate allocations: one plan per plane pair is avoided by
    // addressing everything relative to nullptr and shifting the C offsets for the imaginary pass
    auto gr = plan_gemm_raw(ctx, !left, left, pr);
    std::vector<GemmProblem> pi = pr;
    for (size_t k = 0; k < pi.size(); k++) pi[k].c = off(Ei[which_g[k]]);
    auto gi2 = plan_gemm_raw(ctx, !left, left, pi);
    const double* Mr = nullptr;                  // offsets are absolute addresses / 8
    const double* Mi = Mr + tn;                  // imaginary plane of T: + nelem doubles
    // left : rho = M M^+  -> A = M, B^T = conj(M):  Cr = Mr Mr^T + Mi Mi^T ; Ci = Mi Mr^T - Mr Mi^T
    // right: rho = M^+ M  -> A^T = conj(M), B = M:  Cr = Mr^T Mr + Mi^T Mi ; Ci = Mr^T Mi - Mi^T Mr
    run_gemm(ctx, *gr, Mr, Mr, nullptr, 1.0, false);
    run_gemm(ctx, *gr, Mi, Mi, nullptr, 1.0, true);
    if (left) {
      run_gemm(ctx, *gi2, Mi, Mr, nullptr, 1.0, false);
      run_gemm(ctx, *gi2, Mr, Mi, nullptr, -1.0, true);
    } else {
      run_gemm(ctx, *gi2, Mr, Mi, nullptr, 1.0, false);
      run_gemm(ctx, *gi2, Mi, Mr, nullptr, -1.0, true);
    }
    ctx->sync();
    // rho += noise * X X^+ : the same four products per operand, accumulated with alpha = noise.  The plane distance
    // differs per operand tensor, hence one plan per (operand, group).
    for (size_t t = 0; t < nops.size(); t++) {
      for (size_t gi = 0; gi < fg.size(); gi++) {
        FG& f = fg[gi];
        const int64_t n = left ? f.R : f.C;
        if (n == 0 || f.Xs.size() <= t || !f.Xs[t].X || f.Xs[t].XK == 0) continue;
        GemmProblem p{};
        p.M = p.N = (int)n; p.ldc = (int)n; p.K = (int)f.Xs[t].XK;
        p.a = p.b = off(f.Xs[t].X); p.lda = p.ldb = (int)f.Xs[t].ldx;
        p.c = off(Er[gi]);
        auto gx = plan_gemm_raw(ctx, !left, left, {p});
        GemmProblem q = p;
        q.c = off(Ei[gi]);
        auto gy = plan_gemm_raw(ctx, !left, left, {q});
        const double* Xr = nullptr;
        const double* Xi = Xr + f.Xs[t].xn;
        const double a = prm.noise;
        run_gemm(ctx, *gx, Xr, Xr, nullptr, a, true);
        if (f.Xs[t].xn) {
          run_gemm(ctx, *gx, Xi, Xi, nullptr, a, true);
          if (left) {
            run_gemm(ctx, *gy, Xi, Xr, nullptr, a, true);
            run_gemm(ctx, *gy, Xr, Xi, nullptr, -a, true);
          } else {
            run_gemm(ctx, *gy, Xr, Xi, nullptr, a, true);
            run_gemm(ctx, *gy, Xi, Xr, nullptr, -a, true);
          }
        }
        ctx->sync();
      }
    }
  }
  // ---- Hermitian eigendecompositions: planar -> interleaved, cusolverDnZheevd per charge group on the SM
  // partitions (syevd_batch), eigenvectors back to planes
  std::vector<double> pool;
  {
    std::vector<EighJob> jobs;
    std::vector<size_t> job_group;
    std::vector<double2*> Zs;
    for (size_t gi = 0; gi < fg.size(); gi++) {
      FG& f = fg[gi];
      const int64_t n = left ? f.R : f.C;
      if (n == 0) continue;
      double2* Z = (double2*)talloc(2 * n * n);
      double* Wv = talloc(n);
      const int grid = (int)std::min<int64_t>((n * n + 255) / 256, 1184);
      interleave_kernel<<<grid, 256, 0, ctx->stream>>>(Z, Er[gi], Ei[gi], n * n);
      CUDA_OK(cudaGetLastError());
      EighJob j{n, (double*)Z, Wv, &f.vals, n};
      j.cplx = true;
      jobs.push_back(j);
      job_group.push_back(gi);
      Zs.push_back(Z);
    }
    syevd_batch(ctx, jobs);
    for (size_t k = 0; k < jobs.size(); k++) {
      const size_t gi = job_group[k];
      const int64_t n = jobs[k].n;
      const int grid = (int)std::min<int64_t>((n * n + 255) / 256, 1184);
      deinterleave_kernel<<<grid, 256, 0, ctx->stream>>>(Er[gi], Ei[gi], Zs[k], n * n);     // eigenvectors, planar
      CUDA_OK(cudaGetLastError());
      for (double w : fg[gi].vals) pool.push_back(std::fabs(w));
    }
    ctx->sync();
  }
  // ---- pooled truncation (identical to the real path)
  std::sort(pool.begin(), pool.end(), std::greater<double>());
  FactorizeResult res;
  res.path = "eigen(complex)";
  double docut = 0.0;
  if (!gauge_only) truncate_spectrum(pool, prm.maxdim, prm.mindim, prm.cutoff, res.truncerr, docut);
  res.eigs = pool;
  Index m;
  m.nq = T.inds[0].nq;
  std::vector<size_t> kept;
  for (size_t gi = 0; gi < fg.size(); gi++) {
    FG& f = fg[gi];
    if (f.vals.empty()) continue;
    std::vector<int> order(f.vals.size());
    std::iota(order.begin(), order.end(), 0);
    auto wt = [&](int i) { return std::fabs(f.vals[i]); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return wt(a) > wt(b); });
    if (gauge_only) {
      const int64_t k = std::min(f.R, f.C);
      for (int64_t i = 0; i < k; i++) f.keep.push_back(order[i]);
    } else
    for (int i : order)
      if (wt(i) > docut) f.keep.push_back(i);
    if (f.keep.empty()) continue;
    m.dims.push_back((int)f.keep.size());
    m.qns.push_back(f.q);
    kept.push_back(gi);
  }
  m.dir = prm.new_dir_on_L;
  std::vector<Index> li(T.inds.begin(), T.inds.begin() + split), ri;
  li.push_back(m);
  ri.push_back(m.dag());
  ri.insert(ri.end(), T.inds.begin() + split, T.inds.end());
  res.L = std::make_shared<Tensor>(ctx, li, split, true, true);
  res.R = std::make_shared<Tensor>(ctx, ri, 1, true, true);
  const int64_t ln = res.L->nelem, rn = res.R->nelem;
  std::vector<GemmProblem> proj;
  std::vector<size_t> proj_group;             // (unused bookkeeping of skipped groups keeps proj aligned with `kept`)
  for (size_t gi : kept) {
    FG& f = fg[gi];
    const int nk = (int)f.keep.size();
    const int gl = res.L->find_group(f.q), gr = res.R->find_group(f.q);
    // a sector opened by the noise term alone exists only on the factor that carries the eigenvectors
    TNL_CHECK(left ? gl >= 0 : gr >= 0, "new link sector missing on the isometric factor");
    TNL_CHECK((gl >= 0 && gr >= 0) || !f.M, "new link sector missing on a factor");
    static const Group kNoGroup{};
    const Group &GL = gl >= 0 ? res.L->groups[gl] : kNoGroup, &GR = gr >= 0 ? res.R->groups[gr] : kNoGroup;
    if (gl >= 0) TNL_CHECK(GL.R == f.R && GL.C == nk, "factor group shape");
    if (gr >= 0) TNL_CHECK(GR.R == nk && GR.C == f.C, "factor group shape");
    double* Ld = gl >= 0 ? res.L->d + GL.base : nullptr;
    double* Rd = gr >= 0 ? res.R->d + GR.base : nullptr;
    if (left) {
      // L = V_k ; R = V_k^+ M
      gather_cols(ctx, Ld, GL.ld, Er[gi], f.R, f.R, f.keep, nullptr);
      gather_cols(ctx, Ld + ln, GL.ld, Ei[gi], f.R, f.R, f.keep, nullptr);
      if (!f.M || gr < 0) { proj_group.push_back(gi); proj.push_back(GemmProblem{}); proj.back().M = 0; continue; }
      GemmProblem p{};
      p.M = nk; p.N = (int)f.C; p.K = (int)f.R;
      p.a = off(Ld); p.lda = (int)GL.ld; p.b = off(f.M); p.ldb = (int)f.ldm;
      p.c = off(Rd); p.ldc = (int)GR.ld;
      proj.push_back(p);
    } else {
      // R = V_k^+ (rows) ; L = M V_k
      const int64_t ldk = (f.C + 1) & ~int64_t(1);
      double* Vk = talloc(2 * ldk * nk);          // planes of ldk*nk
      gather_cols(ctx, Vk, ldk, Er[gi], f.C, f.C, f.keep, nullptr);
      gather_cols(ctx, Vk + ldk * nk, ldk, Ei[gi], f.C, f.C, f.keep, nullptr);
      transpose(ctx, Rd, GR.ld, Vk, ldk, f.C, nk);
      transpose(ctx, Rd + rn, GR.ld, Vk + ldk * nk, ldk, f.C, nk);
      std::vector<double> neg(nk, -1.0);
      double* dneg = ctx->upload(neg);
      scale_rows_or_cols(ctx, Rd + rn, GR.ld, nk, f.C, dneg, true);      // conjugate
      ctx->free(dneg);
      if (!f.M || gl < 0) { proj_group.push_back(gi); proj.push_back(GemmProblem{}); proj.back().M = 0; continue; }
      GemmProblem p{};
      p.M = (int)f.R; p.N = nk; p.K = (int)f.C;
      p.a = off(f.M); p.lda = (int)f.ldm; p.b = off(Vk); p.ldb = (int)ldk;
      p.c = off(Ld); p.ldc = (int)GL.ld;
      // the imaginary plane of Vk sits ldk*nk doubles after the real one: remember it through the K-independent offset
      proj.push_back(p);
      f.ldu = ldk * nk;                            // plane distance of this group's Vk
    }
  }
  if (!proj.empty()) {
    if (left) {
      // R = conj(L)^T M : planes of L are ln apart, of M tn apart, of R rn apart
      std::vector<GemmProblem> live;
      for (const GemmProblem& q : proj) if (q.M > 0) live.push_back(q);
      if (!live.empty()) {
        auto g = plan_gemm_raw(ctx, true, false, live);
        const double* z = nullptr;
        cgemm(ctx, *g, z, z + ln, true, z, z + tn, false, (double*)z, (double*)z + rn);
        ctx->sync();
      }
    } else {
      // L = M V_k : the plane distance of Vk differs per group -> one launch set per group
      for (size_t k = 0; k < proj.size(); k++) {
        if (proj[k].M == 0) continue;
        auto g = plan_gemm_raw(ctx, false, false, {proj[k]});
        const double* z = nullptr;
        cgemm(ctx, *g, z, z + tn, false, z, z + fg[kept[k]].ldu, false, (double*)z, (double*)z + ln);
        ctx->sync();
      }
    }
    ctx->sync();
  }
  ctx->sync();
  for (void* p : temps) ctx->free(p);
  return res;
}

// Multi-GPU spectrum exchange: every rank fills the segments of its own groups, the rest is zero; one all-reduce.
// Afterwards every rank holds the values of every group (f.nvals each).
static void exchange_spectrum(Ctx* ctx, std::vector<FG>& fg) {
  int64_t tot = 0;
  for (FG& f : fg) tot += f.nvals;
  std::vector<double> all((size_t)((tot + 1) & ~int64_t(1)), 0.0);
  int64_t o = 0;
  for (FG& f : fg) {
    if (f.mine) for (int64_t i = 0; i < (int64_t)f.vals.size(); i++) all[o + i] = f.vals[i];
    o += f.nvals;
  }
  if (tot > 0) {
    double* d = ctx->upload(all);
    comm_allreduce_sum(ctx, d, (int64_t)all.size());
    CUDA_OK(cudaMemcpyAsync(all.data(), d, all.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d);
  }
  o = 0;
  for (FG& f : fg) {
    f.vals.assign(all.begin() + o, all.begin() + o + f.nvals);
    o += f.nvals;
  }
}

// Deflated Gram refinement -- what makes the Gram-matrix SVD as accurate as a bidiagonalisation-based one.
// The eigenvalues of rho = M M^T come out with an ABSOLUTE error eps * sigma_max^2, so weights below ~1e-10 of the
// largest carry few correct digits (a converged DMRG state keeps weights down to the cutoff, 1e-15).  The eigenvectors
// U, however, are orthonormal to eps and split the space cleanly: the span U_T of the unresolved ones contains the
// small singular directions, and in that subspace the problem is well scaled again,
//     B = U_T^T M   (|T| x C),   rho_2 = B B^T  with  ||rho_2|| <= tau * sigma_max^2,
// so one more eigendecomposition of the |T| x |T| matrix resolves ten further orders of magnitude (leakage of the
// resolved directions into rho_2 is O(eps^2)).  Repeating this at most three times covers the whole double range.
// Only directions that can still matter for the truncation are refined: weights more than 1e3 below the smallest
// kept weight stay as they are (they are discarded either way and enter only the truncation error).  Everything is
// grouped DGEMMs on the DMMA kernel plus eigendecompositions of shrinking size.
template <class TAlloc>
static void refine_gram_spectrum(Ctx* ctx, std::vector<FG>& fg, bool left, const FactorizeParams& prm, TAlloc& talloc) {
  const double tau = 1e-10;
  auto off = [](const double* p) { return (int64_t)(reinterpret_cast<intptr_t>(p) / (intptr_t)sizeof(double)); };
  auto ev = [](int64_t x) { return (x + 1) & ~int64_t(1); };
  for (int level = 1; level <= 3; level++) {
    if (ctx->shard_world() > 1) exchange_spectrum(ctx, fg);
    std::vector<double> pooled;
    for (FG& f : fg)
      for (int i : f.admissible()) pooled.push_back(std::fabs(f.vals[i]));
    if (pooled.empty()) return;
    std::sort(pooled.begin(), pooled.end(), std::greater<double>());
    const double wmax = pooled[0];
    if (!(wmax > 0.0)) return;
    std::vector<double> kept = pooled;
    double te = 0, dc = 0;
    truncate_spectrum(kept, prm.maxdim, prm.mindim, prm.cutoff, te, dc);
    const double pcut = kept.empty() ? 0.0 : kept.back();
    const double hi = std::pow(tau, level) * wmax;              // weights below `hi` are unresolved so far
    if (pcut >= hi) return;                                     // every kept weight is resolved
    const double noise_floor = 1e-16 * std::pow(tau, level - 1) * wmax;
    const double lo = pcut > 1e3 * noise_floor ? 1e-3 * pcut : -1.0;
    struct Sub { FG* f; std::vector<int> idx; int64_t n, t; double *Ut, *B, *G2, *W2, *Un; std::vector<double> mu; };
    std::vector<Sub> subs;
    for (FG& f : fg) {
      if (!f.mine || !f.E || !f.M) continue;
      Sub sb{};
      sb.f = &f;
      sb.n = left ? f.R : f.C;
      for (int i : f.admissible()) {
        const double w = std::fabs(f.vals[i]);
        if (w < hi && w >= lo) sb.idx.push_back(i);
      }
      sb.t = (int64_t)sb.idx.size();
      if (sb.t > 0) subs.push_back(std::move(sb));
    }
    // (no early return on an empty list: under sharding the other ranks may still have groups to refine and the
    // loop head contains a collective)
    std::vector<GemmProblem> p1, p2, p3;
    for (Sub& sb : subs) {
      FG& f = *sb.f;
      const int64_t n = sb.n, t = sb.t, ldn = ev(n), ldt = ev(t);
      sb.Ut = talloc(ldn * t);
      gather_cols(ctx, sb.Ut, ldn, f.E, n, n, sb.idx, nullptr);
      sb.G2 = talloc(ldt * t);
      sb.W2 = talloc(t);
      sb.Un = talloc(ldn * t);
      GemmProblem a{}, b{}, c{};
      if (left) {
        // B[t x C] = U_T^T M ;  G2 = B B^T
        sb.B = talloc(ldt * f.C);
        a.M = (int)t; a.N = (int)f.C; a.K = (int)f.R; a.a = off(sb.Ut); a.lda = (int)ldn; a.b = off(f.M); a.ldb = (int)f.ldm;
        a.c = off(sb.B); a.ldc = (int)ldt;
        b.M = b.N = (int)t; b.K = (int)f.C; b.a = b.b = off(sb.B); b.lda = b.ldb = (int)ldt; b.c = off(sb.G2); b.ldc = (int)ldt;
      } else {
        // B[R x t] = M V_T ;  G2 = B^T B
        const int64_t ldr = ev(f.R);
        sb.B = talloc(ldr * t);
        a.M = (int)f.R; a.N = (int)t; a.K = (int)f.C; a.a = off(f.M); a.lda = (int)f.ldm; a.b = off(sb.Ut); a.ldb = (int)ldn;
        a.c = off(sb.B); a.ldc = (int)ldr;
        b.M = b.N = (int)t; b.K = (int)f.R; b.a = b.b = off(sb.B); b.lda = b.ldb = (int)ldr; b.c = off(sb.G2); b.ldc = (int)ldt;
      }
      // refined vectors  U_T V2  [n x t]
      c.M = (int)n; c.N = (int)t; c.K = (int)t; c.a = off(sb.Ut); c.lda = (int)ldn; c.b = off(sb.G2); c.ldb = (int)ldt;
      c.c = off(sb.Un); c.ldc = (int)ldn;
      p1.push_back(a); p2.push_back(b); p3.push_back(c);
    }
    if (!subs.empty()) {
      auto g1 = plan_gemm_raw(ctx, left, false, p1);             // left: op(A) = U_T^T
      run_gemm(ctx, *g1, nullptr, nullptr, nullptr);
      auto g2 = plan_gemm_raw(ctx, !left, left, p2, true);       // left: B B^T ; right: B^T B (lower triangle)
      run_gemm(ctx, *g2, nullptr, nullptr, nullptr);
      ctx->sync();
      std::vector<EighJob> jobs;
      for (Sub& sb : subs) jobs.push_back(EighJob{sb.t, sb.G2, sb.W2, &sb.mu, ev(sb.t)});
      syevd_batch(ctx, jobs);
      auto g3 = plan_gemm_raw(ctx, false, false, p3);
      run_gemm(ctx, *g3, nullptr, nullptr, nullptr);
      for (Sub& sb : subs) {
        FG& f = *sb.f;
        scatter_cols(ctx, f.E, sb.n, sb.Un, ev(sb.n), sb.n, sb.idx);
        for (int64_t k = 0; k < sb.t; k++) f.vals[sb.idx[k]] = sb.mu[k];
      }
      ctx->sync();
    }
  }
}

FactorizeResult factorize(Ctx* ctx, const Tensor& T, const FactorizeParams& prm) {
  if (T.cplx) return factorize_complex(ctx, T, prm);
  const int split = T.nrow, rank = T.rank();
  TNL_CHECK(split >= 1 && split < rank, "factorize needs a proper bipartition");
  int which = prm.which;
  const bool noisy = prm.noise != 0.0 && prm.noiseX != nullptr;
  if (which == 0) which = noisy ? 2 : (prm.cutoff <= 1e-12 ? 1 : 2);
  const bool left = prm.ortho_left != 0;
  cusolverDnHandle_t H = solver(ctx);
  int svd_alg = prm.svd_alg;
  if (which == 3) return factorize_qr(ctx, T, left, prm.new_dir_on_L);
  // SVD drivers: 0 ("divide_and_conquer", the reference's default) and 2 ("gram") = singular vectors from the
  // Hermitian eigenproblem of M M^T / M^T M built with the grouped DGEMM, followed by the deflated refinement
  // (refine_gram_spectrum) that restores the accuracy of a bidiagonalisation-based SVD for small singular values;
  // 1 = cusolverDnXgesvdp (polar), 3 = cusolverDnDgesvd (QR iteration).
  const bool gram = (which == 1 && (svd_alg == 0 || svd_alg == 2));
  if (gram) which = 2;

  // ---- collect charge groups (union of T's and, on the kept side, the noise operand's)
  std::vector<FG> fg;
  for (const Group& g : T.groups) {
    FG f; f.q = g.q; f.R = g.R; f.C = g.C; f.M = T.d + g.base; f.ldm = g.ld;
    fg.push_back(f);
  }
  std::vector<const Tensor*> nops;
  if (noisy && which == 2) {
    nops.push_back(prm.noiseX);
    for (const Tensor* x : prm.noiseXmore) nops.push_back(x);
    for (size_t t = 0; t < nops.size(); t++) {
      const Tensor& X = *nops[t];
      for (const Group& g : X.groups) {
        // ortho left : X = [(rows) | K] rows match T's row group;  ortho right: X = [K | (cols)] cols match T's cols
        Charge q = g.q;
        auto it = std::find_if(fg.begin(), fg.end(), [&](const FG& f) { return f.q == q; });
        if (it == fg.end()) {
          FG f; f.q = q;
          if (left) { f.R = g.R; f.C = 0; } else { f.R = 0; f.C = g.C; }
          fg.push_back(f);
          it = fg.end() - 1;
        }
        if (left) TNL_CHECK(it->R == g.R, "noise operand rows do not match");
        else TNL_CHECK(it->C == g.C, "noise operand cols do not match");
        it->Xs.resize(nops.size());
        it->Xs[t].X = X.d + g.base; it->Xs[t].ldx = g.ld; it->Xs[t].XK = left ? g.C : g.R;
      }
    }
    std::sort(fg.begin(), fg.end(), [](const FG& a, const FG& b) { return a.q < b.q; });
  }

  // Multi-GPU: the per-charge-group decompositions are independent -- distribute them
  // over the ranks (longest-processing-time first on the measured cost of a decomposition), exchange the spectrum, and sum the two factors.
  const int W = ctx->shard_world();
  if (W > 1) {
    std::vector<size_t> ord(fg.size());
    std::iota(ord.begin(), ord.end(), 0);
    // a rank runs its decompositions concurrently on SM partitions (syevd_batch), so its load is the modelled
    // makespan of its job list, not the sum of the costs; groups go, largest first, to the rank whose makespan
    // grows the least
    auto order_of = [&](const FG& f) { return (int64_t)(which == 1 ? std::max(f.R, f.C) : (left ? f.R : f.C)); };
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return order_of(fg[a]) > order_of(fg[b]); });
    std::vector<std::vector<int64_t>> jobs_of(W);
    std::vector<double> load(W, 0.0);
    for (size_t i : ord) {
      int best = 0;
      double bt = 1e300;
      for (int r = 0; r < W; r++) {
        jobs_of[r].push_back(order_of(fg[i]));
        const double t = eigh_best_makespan(jobs_of[r], ctx->num_sms);
        jobs_of[r].pop_back();
        if (t < bt - 1e-9) { bt = t; best = r; }
      }
      jobs_of[best].push_back(order_of(fg[i]));
      load[best] = bt;
      fg[i].mine = (best == ctx->rank);
    }
  }
  for (FG& f : fg) {
    f.nvals = which == 1 ? std::min(f.R, f.C) : (left ? f.R : f.C);
    if (gram) f.kmax = std::min(f.R, f.C);
  }

  std::vector<void*> temps;                        // overflow allocations (arena too small this time)
  ctx->arena_reset();
  auto talloc = [&](int64_t n) {
    n = (std::max<int64_t>(n, 1) + 1) & ~int64_t(1);
    bool in_arena = false;
    double* p = ctx->arena_alloc((size_t)n, &in_arena);
    CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(double), ctx->stream));
    if (!in_arena) temps.push_back(p);
    return p;
  };
  std::vector<double> pool;

  if (which == 1) {
    // ------------------------------------------------------------------ SVD path
    for (FG& f : fg) {
      const int64_t R = f.R, C = f.C, k = std::min(R, C);
      if (k == 0 || !f.mine) continue;
      const bool tr = R < C;                       // gesvd needs m >= n
      const int64_t m = tr ? C : R, n = tr ? R : C;
      double* Awork = talloc(m * n);
      if (tr) transpose(ctx, Awork, m, f.M, f.ldm, R, C);
      else copy2d(ctx, Awork, m, f.M, f.ldm, R, C);
      double* S = talloc(n);
      double* Uw = talloc(m * n);
      double* Vw = talloc(n * n);
      const int alg = (n >= 64) ? svd_alg : 3;
      if (alg == 1) {
        // polar-decomposition SVD (QDWH + syevd inside cuSOLVER): same absolute accuracy as gesvd, GEMM-rich
        static cusolverDnParams_t params = nullptr;
        if (!params) CUSOLVER_OK(cusolverDnCreateParams(&params));
        size_t wd = 0, wh = 0;
        double* Vfull = talloc(n * n);
        CUSOLVER_OK(cusolverDnXgesvdp_bufferSize(H, params, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, CUDA_R_64F, Awork, m,
                                                 CUDA_R_64F, S, CUDA_R_64F, Uw, m, CUDA_R_64F, Vfull, n, CUDA_R_64F, &wd, &wh));
        double* work = solver_ws(ctx, wd / sizeof(double) + 2);
        std::vector<char> hbuf(wh + 8);
        double herr = 0;
        CUSOLVER_OK(cusolverDnXgesvdp(H, params, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, CUDA_R_64F, Awork, m, CUDA_R_64F, S,
                                      CUDA_R_64F, Uw, m, CUDA_R_64F, Vfull, n, CUDA_R_64F, work, wd, hbuf.data(), wh,
                                      ctx->d_info, &herr));
        transpose(ctx, Vw, n, Vfull, n, n, n);    // Vw = V^T
      } else {
        int lwork = 0;
        CUSOLVER_OK(cusolverDnDgesvd_bufferSize(H, (int)m, (int)n, &lwork));
        double* work = solver_ws(ctx, (size_t)lwork + 2 * n);
        CUSOLVER_OK(cusolverDnDgesvd(H, 'S', 'S', (int)m, (int)n, Awork, (int)m, S, Uw, (int)m, Vw, (int)n, work, lwork,
                                     work + lwork, ctx->d_info));
      }
      f.vals.resize(n);
      CUDA_OK(cudaMemcpyAsync(f.vals.data(), S, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      int info = 0;
      CUDA_OK(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
      TNL_CHECK(info == 0, "cuSOLVER SVD did not converge");
      if (!tr) {
        f.U = Uw; f.ldu = m; f.Vt = Vw; f.ldv = n;
      } else {                                      // M^T = Uw S Vw  =>  M = Vw^T S Uw^T
        f.U = talloc(R * k); f.ldu = R;
        transpose(ctx, f.U, R, Vw, n, n, n);        // U = Vw^T   [R x R]
        f.Vt = talloc(k * C); f.ldv = k;
        transpose(ctx, f.Vt, k, Uw, m, C, k);       // Vt = Uw^T  [k x C]
      }
      for (double s : f.vals) pool.push_back(s * s);
    }
  } else {
    // ------------------------------------------------------------------ eigen path
    std::vector<GemmProblem> pr, prn;
    auto off = [](const double* p) { return (int64_t)(reinterpret_cast<intptr_t>(p) / (intptr_t)sizeof(double)); };
    for (FG& f : fg) {
      const int64_t n = left ? f.R : f.C;
      if (n == 0 || !f.mine) continue;
      f.E = talloc(n * n);
      GemmProblem p{};
      p.M = p.N = (int)n; p.ldc = (int)n; p.c = off(f.E);
      if (f.M) {
        p.K = (int)(left ? f.C : f.R);
        p.a = p.b = off(f.M); p.lda = p.ldb = (int)f.ldm;
      } else {
        p.K = 0;
      }
      pr.push_back(p);
    }
    // left: rho = M M^T  (A = M, B^T = M);  right: rho = M^T M (A^T = M, B = M)
    auto g = plan_gemm_raw(ctx, !left, left, pr, /*lower_only: syevd reads the lower triangle*/ true);
    run_gemm(ctx, *g, nullptr, nullptr, nullptr);
    if (noisy) {
      // rho += noise * sum_t X_t X_t^T (one grouped DGEMM + axpy per MPO term; the D buffers are reused)
      std::vector<double*> dr(fg.size(), nullptr);
      for (size_t t = 0; t < nops.size(); t++) {
        prn.clear();
        std::vector<char> has(fg.size(), 0);
        size_t i = 0;
        for (FG& f : fg) {
          const int64_t n = left ? f.R : f.C;
          const size_t gi = i++;
          if (n == 0 || !f.mine || f.Xs.size() <= t || !f.Xs[t].X || f.Xs[t].XK == 0) continue;
          if (!dr[gi]) dr[gi] = talloc(n * n);
          has[gi] = 1;
          GemmProblem p{};
          p.M = p.N = (int)n; p.K = (int)f.Xs[t].XK; p.ldc = (int)n; p.c = off(dr[gi]);
          p.a = p.b = off(f.Xs[t].X); p.lda = p.ldb = (int)f.Xs[t].ldx;
          prn.push_back(p);
        }
        if (prn.empty()) continue;
        auto gn = plan_gemm_raw(ctx, !left, left, prn, true);
        run_gemm(ctx, *gn, nullptr, nullptr, nullptr);
        i = 0;
        for (FG& f : fg) {
          const int64_t n = left ? f.R : f.C;
          const size_t gi = i++;
          // talloc buffers are zero-initialised and padded to an even length, so the padded axpy is exact
          if (has[gi]) vec_axpy(ctx, f.E, dr[gi], (n * n + 1) & ~int64_t(1), prm.noise);
        }
        ctx->sync();                               // gn's device arrays die with this iteration
      }
    }
    // Hermitian eigendecompositions, one per charge group
    std::vector<EighJob> jobs;
    for (FG& f : fg) {
      const int64_t n = left ? f.R : f.C;
      if (n == 0 || !f.mine) continue;
      jobs.push_back(EighJob{n, f.E, talloc(n), &f.vals, n});
    }
    ctx->sync();
    g.reset();
    syevd_batch(ctx, jobs);
    if (gram) refine_gram_spectrum(ctx, fg, left, prm, talloc);
    for (FG& f : fg)
      for (int i : f.admissible()) pool.push_back(std::fabs(f.vals[i]));
  }

  if (W > 1) {
    exchange_spectrum(ctx, fg);
    pool.clear();
    for (FG& f : fg)
      for (int i : f.admissible()) pool.push_back(which == 1 ? f.vals[i] * f.vals[i] : std::fabs(f.vals[i]));
  }

  // ---- pooled truncation
  std::sort(pool.begin(), pool.end(), std::greater<double>());
  FactorizeResult res;
  res.path = which == 1 ? "svd" : (gram ? "svd(gram, deflated refinement)" : "eigen");
  double docut = 0.0;
  truncate_spectrum(pool, prm.maxdim, prm.mindim, prm.cutoff, res.truncerr, docut);
  res.eigs = pool;
  Index m;
  m.nq = T.inds[0].nq;
  std::vector<FG*> kept;
  for (FG& f : fg) {
    if (f.vals.empty()) continue;
    auto wt = [&](int i) { return which == 1 ? f.vals[i] * f.vals[i] : std::fabs(f.vals[i]); };
    for (int i : f.admissible())
      if (wt(i) > docut) f.keep.push_back(i);
    if (f.keep.empty()) continue;
    m.dims.push_back((int)f.keep.size());
    m.qns.push_back(f.q);
    kept.push_back(&f);
  }
  m.dir = prm.new_dir_on_L;
  std::vector<Index> li(T.inds.begin(), T.inds.begin() + split), ri;
  li.push_back(m);
  ri.push_back(m.dag());
  ri.insert(ri.end(), T.inds.begin() + split, T.inds.end());
  res.L = std::make_shared<Tensor>(ctx, li, split);
  res.R = std::make_shared<Tensor>(ctx, ri, 1);

  std::vector<GemmProblem> proj;     // eigen path: the other factor by projection
  auto off = [](const double* p) { return (int64_t)(reinterpret_cast<intptr_t>(p) / (intptr_t)sizeof(double)); };
  for (FG* fp : kept) {
    FG& f = *fp;
    const int nk = (int)f.keep.size();
    int gl = res.L->find_group(f.q);
    int gr = res.R->find_group(f.q);
    // a sector opened by the noise term alone exists only on the factor that carries the eigenvectors
    TNL_CHECK(gl >= 0 || (which == 2 && !left), "new link sector missing on the left factor");
    static const Group kNoGroup{};
    const Group& GL = gl >= 0 ? res.L->groups[gl] : kNoGroup;
    if (gl >= 0) TNL_CHECK(GL.R == f.R && GL.C == nk, "left factor group shape");
    double* Ld = gl >= 0 ? res.L->d + GL.base : nullptr;
    if (!f.mine) continue;                         // another rank fills this group; summed by the all-reduce below
    if (which == 1) {
      TNL_CHECK(gr >= 0, "new link sector missing on the right factor");
      const Group& GR = res.R->groups[gr];
      TNL_CHECK(GR.R == nk && GR.C == f.C, "right factor group shape");
      double* Rd = res.R->d + GR.base;
      std::vector<double> sv(nk);
      for (int i = 0; i < nk; i++) sv[i] = f.vals[f.keep[i]];
      // gesvd returns sigma descending, so keep = 0..nk-1: plain 2D copies
      bool ident = true;
      for (int i = 0; i < nk; i++) ident = ident && f.keep[i] == i;
      TNL_CHECK(ident, "gesvd spectrum is expected in descending order");
      copy2d(ctx, Ld, GL.ld, f.U, f.ldu, f.R, nk);
      copy2d(ctx, Rd, GR.ld, f.Vt, f.ldv, nk, f.C);
      double* dsv = ctx->upload(sv);
      if (left) scale_rows_or_cols(ctx, Rd, GR.ld, nk, f.C, dsv, true);
      else scale_rows_or_cols(ctx, Ld, GL.ld, f.R, nk, dsv, false);
      ctx->free(dsv);
    } else if (left) {
      gather_cols(ctx, Ld, GL.ld, f.E, f.R, f.R, f.keep, nullptr);
      if (gr >= 0 && f.M) {                       // R = U^T M
        const Group& GR = res.R->groups[gr];
        TNL_CHECK(GR.R == nk && GR.C == f.C, "right factor group shape");
        GemmProblem p{};
        p.M = nk; p.N = (int)f.C; p.K = (int)f.R;
        p.a = off(Ld); p.lda = (int)GL.ld; p.b = off(f.M); p.ldb = (int)f.ldm;
        p.c = off(res.R->d + GR.base); p.ldc = (int)GR.ld;
        proj.push_back(p);
      }
    } else {
      // eigenvectors of M^T M are the rows of the right factor
      TNL_CHECK(gr >= 0, "new link sector missing on the right factor");
      const Group& GR = res.R->groups[gr];
      TNL_CHECK(GR.R == nk && GR.C == f.C, "right factor group shape");
      const int64_t ldk = (f.C + 1) & ~int64_t(1);      // even leading dimension: Vk is a GEMM operand below
      double* Vk = talloc(ldk * nk);
      gather_cols(ctx, Vk, ldk, f.E, f.C, f.C, f.keep, nullptr);
      transpose(ctx, res.R->d + GR.base, GR.ld, Vk, ldk, f.C, nk);
      if (f.M && gl >= 0) {                       // L = M V
        GemmProblem p{};
        p.M = (int)f.R; p.N = nk; p.K = (int)f.C;
        p.a = off(f.M); p.lda = (int)f.ldm; p.b = off(Vk); p.ldb = (int)ldk;
        p.c = off(Ld); p.ldc = (int)GL.ld;
        proj.push_back(p);
      }
    }
  }
  if (!proj.empty()) {
    auto g = plan_gemm_raw(ctx, which == 2 && left, false, proj);
    run_gemm(ctx, *g, nullptr, nullptr, nullptr);
    ctx->sync();
  }
  if (W > 1) {
    comm_allreduce_sum(ctx, res.L->d, res.L->nelem);
    comm_allreduce_sum(ctx, res.R->d, res.R->nelem);
  }
  ctx->sync();
  for (void* p : temps) ctx->free(p);
  return res;
}

}  // namespace tnl
