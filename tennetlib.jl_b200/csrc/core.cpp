// Host-side planner of tnl_b200: charge-fused layouts, grouped-GEMM plans and transform plans.
// See core.hpp for the data model.  The block-matching rule restated here is NDTensors'
// `contract_blockoffsets` (blocks pair when their contracted sector numbers agree), specialised to
// the fused layout where all pairs of one charge sector collapse into a single dense GEMM.
#include "core.hpp"

#include <algorithm>
#include <cstring>
#include <functional>

namespace tnl {

// ------------------------------------------------------------------------------------------ Ctx
Ctx::Ctx(int dev) : device(dev) {
  CUDA_OK(cudaSetDevice(dev));
  CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  num_sms = prop.multiProcessorCount;
  cudaMemPool_t pool;
  CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t thr = UINT64_MAX;
  CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  CUDA_OK(cudaMalloc(&d_scalars, 256 * sizeof(double)));
  CUDA_OK(cudaMallocHost(&h_scalars, 256 * sizeof(double)));
  CUDA_OK(cudaMalloc(&d_sync, sizeof(unsigned int)));
  CUDA_OK(cudaMemset(d_sync, 0, sizeof(unsigned int)));
  CUDA_OK(cudaMalloc(&d_partials, 4096 * sizeof(double)));
  CUDA_OK(cudaMalloc(&d_info, 16 * sizeof(int)));
  if (const char* e = getenv("TNL_GEMM_TMA")) use_tma = atoi(e) != 0;
  if (const char* e = getenv("TNL_GEMM_DUAL")) dual_gemm = atoi(e) != 0;
}
// Pinned staging memory for cudaMemcpyAsync sources that the host fills and forgets: a ring; when it wraps the
// stream is drained once so that no pending copy still reads the bytes about to be overwritten.
void* Ctx::stage_pinned(size_t bytes) {
  bytes = (bytes + 127) & ~size_t(127);
  if (!pin_ring) {
    pin_cap = size_t(4) << 20;
    CUDA_OK(cudaMallocHost((void**)&pin_ring, pin_cap));
  }
  TNL_CHECK(bytes <= pin_cap, "staging request larger than the pinned ring");
  if (pin_off + bytes > pin_cap) { sync(); pin_off = 0; }
  void* p = pin_ring + pin_off;
  pin_off += bytes;
  return p;
}
Ctx::~Ctx() {
  cudaStreamSynchronize(stream);

  if (solver_work) cudaFree(solver_work);
  for (auto& sl : slots) if (sl.p) cudaFreeAsync(sl.p, stream);
  for (auto& v : vec_pool) cudaFreeAsync(v.first, stream);
  cudaStreamSynchronize(stream);
  cudaFree(d_scalars);
  if (pin_ring) cudaFreeHost(pin_ring);
  cudaFreeHost(h_scalars);
  cudaFree(d_sync);
  cudaFree(d_partials);
  cudaFree(d_info);
  cudaStreamDestroy(stream);
}
// Stream-ordered allocation from the device's default memory pool (release threshold = never): blocks freed on
// the context stream are reused by later allocations of any size without touching the OS.
void* Ctx::alloc(size_t bytes) {
  void* p = nullptr;
  // + 128 bytes: the chunked TMA view of a ragged GEMM operand may read up to 15 doubles past its last row
  // (kernels.cu, gemm_tma_ws_kernel); the slack keeps that over-read inside the allocation
  // Size classes: tensors of neighbouring bonds differ by a few per cent in size; with exact sizes a freed block is
  // slightly too small for its successor, the pool fragments and every so often has to map fresh physical memory in
  // the middle of a sweep (one-off 200 ms stalls in make_phi / position, profiles/r02m_*: step_ms_list).  Rounding
  // large requests up to 16 MB (2 MB for medium ones) lets the successor reuse the block.
  bytes += 128;
  const size_t g = bytes >= (size_t(64) << 20) ? (size_t(16) << 20) : bytes >= (size_t(2) << 20) ? (size_t(2) << 20) : 256;
  bytes = (bytes + g - 1) / g * g;
  CUDA_OK(cudaMallocAsync(&p, bytes, stream));
  return p;
}
void Ctx::free(void* p) {
  if (p) cudaFreeAsync(p, stream);
}

double* Ctx::scratch(int slot, size_t nelem, bool zero) {
  Slot& s = slots[slot];
  size_t bytes = std::max<size_t>(nelem, 2) * sizeof(double);
  if (bytes > s.cap) {
    if (s.p) cudaFreeAsync(s.p, stream);
    slot_epoch++;
    s.cap = bytes + bytes / 4 + 128;                 // headroom: sector sizes drift from bond to bond (+ TMA over-read slack)
    CUDA_OK(cudaMallocAsync(&s.p, s.cap, stream));
  }
  if (zero) CUDA_OK(cudaMemsetAsync(s.p, 0, bytes, stream));
  return (double*)s.p;
}
double* Ctx::vec_acquire(size_t nelem) {
  double* p = nullptr;
  size_t cap = 0;
  for (size_t i = 0; i < vec_pool.size(); i++)
    if (vec_pool[i].second >= nelem) { p = vec_pool[i].first; cap = vec_pool[i].second; vec_pool.erase(vec_pool.begin() + i); break; }
  if (!p) {
    if (!vec_pool.empty()) {                         // recycle the memory of a too-small buffer
      cudaFreeAsync(vec_pool.back().first, stream);
      vec_pool.pop_back();
    }
    cap = nelem + nelem / 4 + 16;
    CUDA_OK(cudaMallocAsync(&p, cap * sizeof(double), stream));
  }
  CUDA_OK(cudaMemsetAsync(p, 0, nelem * sizeof(double), stream));
  vec_caps[p] = cap;
  return p;
}
void Ctx::vec_release(double* p, size_t) {
  if (!p) return;
  auto it = vec_caps.find(p);
  if (it == vec_caps.end()) { cudaFreeAsync(p, stream); return; }
  vec_pool.emplace_back(p, it->second);
  vec_caps.erase(it);
}
void Ctx::arena_reset() {
  if (arena_want > slots[SLOT_ARENA].cap / sizeof(double)) scratch(SLOT_ARENA, arena_want, false);
  arena_off = 0;
  arena_want = 0;
}
double* Ctx::arena_alloc(size_t nelem, bool* from_arena) {
  nelem = (nelem + 31) & ~size_t(31);                // 256-byte granules
  arena_want += nelem;
  Slot& s = slots[SLOT_ARENA];
  if ((arena_off + nelem) * sizeof(double) <= s.cap) {
    double* p = (double*)s.p + arena_off;
    arena_off += nelem;
    *from_arena = true;
    return p;
  }
  *from_arena = false;
  return (double*)alloc(nelem * sizeof(double));
}

GemmPlan::~GemmPlan() {
  if (ctx) {
    ctx->free(d_probs); ctx->free(d_tiles_big); ctx->free(d_tiles_small); ctx->free(d_tiles_tma); ctx->free(d_maps);
    ctx->free(d_splits); ctx->free(d_splitws);
  }
}
TransformPlan::~TransformPlan() {
  if (ctx) { ctx->free(d_blocks); ctx->free(d_flats); ctx->free(d_groups); }
}

// --------------------------------------------------------------------------------------- Tensor
static void enumerate_combos(const std::vector<Index>& inds, int lo, int hi,
                             std::map<Charge, std::vector<Combo>>& out) {
  // product sectors of inds[lo..hi), first index fastest, grouped by charge sum_i dir_i*qn_i
  int n = hi - lo;
  std::vector<int> c(n, 0);
  if (n == 0) {
    Combo cb{};
    cb.dim = 1; cb.off = 0;
    out[charge_zero()].push_back(cb);
    return;
  }
  for (int i = lo; i < hi; i++)
    if (inds[i].nsect() == 0) return;
  while (true) {
    Combo cb{};
    Charge q = charge_zero();
    cb.dim = 1;
    for (int k = 0; k < n; k++) {
      const Index& ix = inds[lo + k];
      cb.c[k] = c[k];
      cb.d[k] = ix.dims[c[k]];
      cb.dim *= cb.d[k];
      for (int a = 0; a < ix.nq; a++) q[a] += ix.dir * ix.qns[c[k]][a];
    }
    out[q].push_back(cb);
    int k = 0;
    while (k < n) {
      if (++c[k] < inds[lo + k].nsect()) break;
      c[k] = 0;
      k++;
    }
    if (k == n) break;
  }
}

void build_layout(Tensor& t) {
  const int r = t.rank();
  TNL_CHECK(r <= MAXR, "tensor rank too large");
  TNL_CHECK(t.nrow >= 0 && t.nrow <= r, "bad nrow");
  std::map<Charge, std::vector<Combo>> rows, cols;
  enumerate_combos(t.inds, 0, t.nrow, rows);
  enumerate_combos(t.inds, t.nrow, r, cols);
  t.groups.clear(); t.blocks.clear(); t.lut.clear();
  int64_t base = 0;
  for (auto& kv : rows) {                              // std::map: ascending charge
    auto it = cols.find(charge_neg(kv.first));
    if (it == cols.end()) continue;
    Group g;
    g.q = kv.first;
    g.rows = kv.second;
    g.cols = it->second;
    for (auto& cb : g.rows) { cb.off = g.R; g.R += cb.dim; }
    for (auto& cb : g.cols) { cb.off = g.C; g.C += cb.dim; }
    g.ld = (g.R + 1) & ~int64_t(1);
    g.base = base;
    base += g.ld * g.C;
    base = (base + 1) & ~int64_t(1);
    int gi = (int)t.groups.size();
    for (auto& cc : g.cols)
      for (auto& rc : g.rows) {
        Block b{};
        int64_t st = 1;
        for (int k = 0; k < t.nrow; k++) { b.c[k] = rc.c[k]; b.d[k] = rc.d[k]; b.st[k] = st; st *= rc.d[k]; }
        st = g.ld;
        for (int k = t.nrow; k < r; k++) { b.c[k] = cc.c[k - t.nrow]; b.d[k] = cc.d[k - t.nrow]; b.st[k] = st; st *= b.d[k]; }
        b.off = g.base + rc.off + g.ld * cc.off;
        b.group = gi;
        t.lut[Tensor::key(b.c, r)] = (int)t.blocks.size();
        t.blocks.push_back(b);
      }
    t.groups.push_back(std::move(g));
  }
  t.nelem = std::max<int64_t>(base, 2);
}

Tensor::Tensor(Ctx* c, std::vector<Index> ii, int nr, bool alloc, bool cx) : ctx(c), inds(std::move(ii)), nrow(nr), cplx(cx) {
  build_layout(*this);
  if (alloc) {
    d = (double*)ctx->alloc(planes() * nelem * sizeof(double));
    CUDA_OK(cudaMemsetAsync(d, 0, planes() * nelem * sizeof(double), ctx->stream));
  }
}
Tensor::~Tensor() {
  if (d && ctx && owns) ctx->free(d);
}
void Tensor::zero() { CUDA_OK(cudaMemsetAsync(d, 0, planes() * nelem * sizeof(double), ctx->stream)); }

// ------------------------------------------------------------------------------------ GEMM plan
static bool combos_equal(const std::vector<Combo>& a, int na, const std::vector<Combo>& b, int nb) {
  if (a.size() != b.size() || na != nb) return false;
  for (size_t i = 0; i < a.size(); i++) {
    if (a[i].dim != b[i].dim) return false;
    for (int k = 0; k < na; k++)
      if (a[i].c[k] != b[i].c[k] || a[i].d[k] != b[i].d[k]) return false;
  }
  return true;
}

static void make_tiles(GemmPlan& plan) {
  // biggest problems first; tiles of one problem are emitted m-fastest so that concurrently resident
  // CTAs share the B panel and stream A through L2
  std::vector<int> order(plan.probs.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    const auto &x = plan.probs[a], &y = plan.probs[b];
    return (double)x.M * x.N * x.K > (double)y.M * y.N * y.K;
  });
  const bool tma = plan.ctx->use_tma;
  for (int pi : order) {
    const auto& p = plan.probs[pi];
    bool big = p.M > 64 && p.N > 64;
    if (big && tma) {
      // Persistent TMA kernel: CTA b works on tiles b, b + #SMs, ... so the tiles in flight are ~#SMs consecutive
      // list entries.  Super-columns of GN n-tiles, m fastest across them: a wave covers a ~12 x 12 block of the
      // output and touches 12 + 12 operand panels instead of all of A (L2 re-reads across waves).
      const int mt = (p.M + 127) / 128, nt = (p.N + 127) / 128, kt = (p.K + 15) / 16;
      const int GN = 12;
      for (int nb = 0; nb < nt; nb += GN)
        for (int mi = 0; mi < mt; mi++)
          for (int ni = nb; ni < std::min(nt, nb + GN); ni++) {
            if (plan.lower_only && ni > mi) {        // tile strictly above the diagonal of a symmetric product
              plan.flops -= 2.0 * std::min(128, p.M - mi * 128) * (double)std::min(128, p.N - ni * 128) * p.K;
              continue;
            }
            plan.tiles_tma.push_back(TmaTile{p.c, mi * 128, ni * 128, p.M, p.N, kt, p.ldc, pi, 0, 0, 0});
          }
      continue;
    }
    int bm = big ? 128 : 64, bn = big ? 128 : 64;
    auto& tl = big ? plan.tiles_big : plan.tiles_small;
    for (int n0 = 0; n0 < p.N; n0 += bn)
      for (int m0 = 0; m0 < p.M; m0 += bm) {
        if (plan.lower_only && n0 >= m0 + bm) {
          plan.flops -= 2.0 * std::min(bm, p.M - m0) * (double)std::min(bn, p.N - n0) * p.K;
          continue;
        }
        tl.push_back(GemmTile{pi, m0, n0});
      }
  }
  // Split-K: a plan that cannot fill the machine with output tiles but has a long contracted range (environment
  // updates of tree nodes: M = N = chi, K = chi^2; Gram matrices of skinny blocks) runs at tiles / #SMs of the DMMA
  // rate -- 1 TFLOP/s for 256 x 256 x 65536.  Every 128x128 tile is cut into S parts along K (at least 32 k-tiles
  // each) so that about two waves of tiles exist; the parts go to a workspace and are summed in order afterwards.
  const int sms = plan.ctx->num_sms;
  if (!plan.tiles_tma.empty() && (int)plan.tiles_tma.size() < sms) {
    int ktmax = 0;
    for (auto& t : plan.tiles_tma) ktmax = std::max(ktmax, t.ktiles);
    const int S0 = (2 * sms + (int)plan.tiles_tma.size() - 1) / (int)plan.tiles_tma.size();
    if (ktmax >= 64 && S0 >= 2) {
      std::vector<TmaTile> cut;
      std::map<int, int> split_of;                   // problem -> index in plan.splits
      int64_t ws = 0;
      for (const TmaTile& t : plan.tiles_tma) {
        const int S = std::max(1, std::min(S0, t.ktiles / 32));
        if (S < 2) { cut.push_back(t); continue; }
        const auto& p = plan.probs[t.prob];
        auto it = split_of.find(t.prob);
        if (it == split_of.end()) {
          const int ldw = (p.M + 1) & ~1;
          plan.splits.push_back(SplitDesc{p.c, ws, p.M, p.N, p.ldc, ldw, S, 0});
          it = split_of.emplace(t.prob, (int)plan.splits.size() - 1).first;
          ws += (int64_t)S * ldw * p.N;
        }
        const SplitDesc& sd = plan.splits[it->second];
        const int per = (t.ktiles + sd.S - 1) / sd.S;
        for (int s = 0; s < sd.S; s++) {
          TmaTile u = t;
          u.kt0 = s * per;
          u.ktiles = std::max(0, std::min(per, t.ktiles - u.kt0));
          u.c = sd.ws + (int64_t)s * sd.ldw * p.N;
          u.ldc = sd.ldw;
          u.split = 1;
          if (u.ktiles == 0) { u.ktiles = 1; u.kt0 = t.ktiles; }   // beyond K: the TMA zero-fills, the part is zero
          cut.push_back(u);
        }
      }
      if (!plan.splits.empty()) {
        plan.tiles_tma.swap(cut);
        plan.d_splitws = (double*)plan.ctx->alloc((size_t)ws * sizeof(double));
        // (parts of tiles that are not computed -- lower_only -- must not feed garbage into the reduction)
        if (plan.lower_only) CUDA_OK(cudaMemsetAsync(plan.d_splitws, 0, (size_t)ws * sizeof(double), plan.ctx->stream));
        plan.d_splits = plan.ctx->upload(plan.splits);
      }
    }
  }
  plan.d_probs = plan.ctx->upload(plan.probs);
  plan.d_tiles_big = plan.ctx->upload(plan.tiles_big);
  plan.d_tiles_small = plan.ctx->upload(plan.tiles_small);
  plan.d_tiles_tma = plan.ctx->upload(plan.tiles_tma);
  if (!plan.tiles_tma.empty())
    plan.d_maps = plan.ctx->alloc((size_t)GemmPlan::MAPSETS_MAX * 2 * plan.probs.size() * 128);
}

// Re-tile a plan so that EVERY problem runs on the persistent TMA kernel (the fused GEMM -> reduce-scatter needs a
// single kernel with the scatter epilogue; small sectors become partially filled 128x128 tiles, zero-filled by the TMA)
void gemm_plan_force_tma(GemmPlan& plan) {
  TNL_CHECK(plan.splits.empty(), "a split-K plan cannot take the scatter epilogue");
  if (plan.tiles_big.empty() && plan.tiles_small.empty()) return;
  Ctx* ctx = plan.ctx;
  std::vector<char> have(plan.probs.size(), 0);
  for (auto& t : plan.tiles_tma) have[t.prob] = 1;
  for (size_t pi = 0; pi < plan.probs.size(); pi++) {
    if (have[pi]) continue;
    const auto& p = plan.probs[pi];
    const int kt = (p.K + 15) / 16;
    for (int n0 = 0; n0 < p.N; n0 += 128)
      for (int m0 = 0; m0 < p.M; m0 += 128) plan.tiles_tma.push_back(TmaTile{p.c, m0, n0, p.M, p.N, kt, p.ldc, (int)pi, 0, 0, 0});
  }
  plan.tiles_big.clear();
  plan.tiles_small.clear();
  ctx->free(plan.d_tiles_big); ctx->free(plan.d_tiles_small); ctx->free(plan.d_tiles_tma);
  plan.d_tiles_big = nullptr; plan.d_tiles_small = nullptr;
  plan.d_tiles_tma = ctx->upload(plan.tiles_tma);
  if (!plan.d_maps) plan.d_maps = ctx->alloc((size_t)GemmPlan::MAPSETS_MAX * 2 * plan.probs.size() * 128);
  plan.mapsets.clear();
}

std::unique_ptr<GemmPlan> plan_gemm_raw(Ctx* ctx, bool transA, bool transB, const std::vector<GemmProblem>& probs, bool lower_only) {
  auto plan = std::make_unique<GemmPlan>();
  plan->ctx = ctx;
  plan->lower_only = lower_only;
  if (lower_only) for (auto& p : probs) TNL_CHECK(p.M == p.N, "lower_only needs square (symmetric) products");
  plan->transA = transA;
  plan->transB = transB;
  for (auto& p : probs) {
    if (p.M == 0 || p.N == 0) continue;
    if (p.K == 0) {
      for (int n = 0; n < p.N; n++) plan->zero_fill.emplace_back(p.c + (int64_t)n * p.ldc, (int64_t)p.M);
      continue;
    }
    plan->flops += 2.0 * p.M * (double)p.N * p.K;
    plan->probs.push_back(p);
  }
  make_tiles(*plan);
  return plan;
}

std::unique_ptr<GemmPlan> plan_gemm(const Tensor& A, bool transA, const Tensor& B, bool transB, Tensor& C,
                                    bool dagA, bool dagB) {
  auto sgnA = [&](Charge q) { return dagA ? charge_neg(q) : q; };
  auto sgnB = [&](Charge q) { return dagB ? charge_neg(q) : q; };
  auto plan = std::make_unique<GemmPlan>();
  plan->ctx = C.ctx;
  plan->transA = transA;
  plan->transB = transB;
  const int nMa = transA ? A.rank() - A.nrow : A.nrow;
  const int nKa = A.rank() - nMa;
  const int nNb = transB ? B.nrow : B.rank() - B.nrow;
  const int nKb = B.rank() - nNb;
  TNL_CHECK(nKa == nKb, "contracted index groups differ in size");
  TNL_CHECK(C.nrow == nMa && C.rank() - C.nrow == nNb, "output bipartition mismatch");
  for (size_t gc = 0; gc < C.groups.size(); gc++) {
    Group& G = C.groups[gc];
    // A group whose M side has charge q
    int ga = A.find_group(sgnA(transA ? charge_neg(G.q) : G.q));
    bool ok = ga >= 0;
    int gb = -1;
    if (ok) {
      const Group& GA = A.groups[ga];
      // charge of the K combos as seen from A
      Charge kA = transA ? sgnA(GA.q) : charge_neg(sgnA(GA.q));
      // seen from B the same sectors carry the opposite arrows
      Charge kB = charge_neg(kA);
      gb = B.find_group(sgnB(transB ? charge_neg(kB) : kB));
      ok = gb >= 0;
    }
    if (!ok) {
      plan->zero_fill.emplace_back(G.base, G.ld * G.C);
      continue;
    }
    const Group& GA = A.groups[ga];
    const Group& GB = B.groups[gb];
    const auto& Mc = transA ? GA.cols : GA.rows;
    const auto& Ka = transA ? GA.rows : GA.cols;
    const auto& Kb = transB ? GB.cols : GB.rows;
    const auto& Nc = transB ? GB.rows : GB.cols;
    TNL_CHECK(combos_equal(Ka, nKa, Kb, nKb), "contracted sectors of A and B do not line up");
    TNL_CHECK(combos_equal(Mc, nMa, G.rows, C.nrow), "row sectors of A and C do not line up");
    TNL_CHECK(combos_equal(Nc, nNb, G.cols, C.rank() - C.nrow), "column sectors of B and C do not line up");
    GemmProblem p;
    p.a = GA.base;
    p.b = GB.base;
    p.c = G.base;
    p.M = (int)G.R;
    p.N = (int)G.C;
    p.K = (int)(transA ? GA.R : GA.C);
    p.lda = (int)GA.ld;
    p.ldb = (int)GB.ld;
    p.ldc = (int)G.ld;
    if (p.M == 0 || p.N == 0) continue;
    if (p.K == 0) { plan->zero_fill.emplace_back(G.base, G.ld * G.C); continue; }
    plan->flops += 2.0 * p.M * (double)p.N * p.K;
    plan->probs.push_back(p);
  }
  make_tiles(*plan);
  return plan;
}

// ------------------------------------------------------------------------------- transform plan
std::unique_ptr<TransformPlan> plan_transform(const Tensor& X, Tensor& Y, const std::vector<int>& xmap,
                                              const Tensor* W, const std::vector<int>& kpos, const SliceMap* slice) {
  auto plan = std::make_unique<TransformPlan>();
  plan->ctx = Y.ctx;
  const int ry = Y.rank(), rx = X.rank();
  TNL_CHECK((int)xmap.size() == ry, "xmap size");
  TNL_CHECK(xmap[0] == 0, "leading index must be shared and leading in both tensors");
  std::vector<int> newpos, passY;          // Y positions of new / passive (non-leading) indices
  for (int j = 1; j < ry; j++) (xmap[j] < 0 ? newpos : passY).push_back(j);
  const int nk = (int)kpos.size(), nn = (int)newpos.size();
  TNL_CHECK(nk <= 2 && nn <= 2, "at most two contracted / new indices");
  TNL_CHECK((int)passY.size() <= MAXP, "too many passive indices");
  TNL_CHECK((W == nullptr) == (nk == 0 && nn == 0), "W must be given iff indices are contracted");
  TNL_CHECK(rx == 1 + (int)passY.size() + nk, "index bookkeeping mismatch");
  if (W) TNL_CHECK(W->rank() == nk + nn && W->nrow == W->rank(), "W must be (contracted..., new...) in natural layout");
  // W blocks grouped by their new-sector coordinates
  int64_t col = 0;
  std::vector<std::vector<int>> keys;              // (leading sector, passive sectors) of every output block
  for (const Block& yb : Y.blocks) {
    {
      std::vector<int> key{yb.c[0]};
      for (int j : passY) key.push_back(yb.c[j]);
      keys.push_back(std::move(key));
    }
    XfBlock xb{};
    xb.yoff = yb.off;
    xb.I = yb.d[0];
    xb.nd0 = xb.nd1 = 1;
    xb.yns[0] = xb.yns[1] = 0;
    if (nn > 0) { xb.nd0 = yb.d[newpos[0]]; xb.yns[0] = yb.st[newpos[0]]; }
    if (nn > 1) { xb.nd1 = yb.d[newpos[1]]; xb.yns[1] = yb.st[newpos[1]]; }
    int64_t ncol = 1;                                   // passive columns; the new-index outputs share a warp
    const int64_t nnew = (int64_t)xb.nd0 * xb.nd1;
    for (int k = 0; k < MAXP; k++) { xb.pd[k] = 1; xb.yps[k] = 0; }
    for (size_t k = 0; k < passY.size(); k++) {
      xb.pd[k] = yb.d[passY[k]];
      xb.yps[k] = yb.st[passY[k]];
      ncol *= xb.pd[k];
    }
    xb.cbeg = (int)plan->contribs.size();
    // candidate contributions: every W block whose new sectors match (or the single identity one)
    int xc[MAXR];
    xc[0] = yb.c[0];
    for (size_t k = 0; k < passY.size(); k++) xc[xmap[passY[k]]] = yb.c[passY[k]];
    int64_t slice_start = 0;
    if (slice) {
      // the sliced index is passive, or the shared leading index (then the local rows are a sub-range of the source rows)
      TNL_CHECK(slice->ypos >= 0 && xmap[slice->ypos] >= 0, "sliced index must not be contracted");
      xc[xmap[slice->ypos]] = slice->orig[yb.c[slice->ypos]];
      slice_start = slice->start[yb.c[slice->ypos]];
    }
    auto add = [&](const Block* wb) {
      for (int k = 0; k < nk; k++) xc[kpos[k]] = wb->c[k];
      int bi = X.find(xc);
      if (bi < 0) return;
      const Block& sb = X.blocks[bi];
      XfContrib c{};
      c.xoff = sb.off + (slice ? slice_start * sb.st[xmap[slice->ypos]] : 0);
      for (int k = 0; k < MAXP; k++) c.xps[k] = 0;
      for (size_t k = 0; k < passY.size(); k++) c.xps[k] = sb.st[xmap[passY[k]]];
      c.kd0 = c.kd1 = 1; c.ks0 = c.ks1 = 0;
      if (nk > 0) { c.kd0 = sb.d[kpos[0]]; c.ks0 = sb.st[kpos[0]]; }
      if (nk > 1) { c.kd1 = sb.d[kpos[1]]; c.ks1 = sb.st[kpos[1]]; }
      c.woff = wb ? wb->off : 0;
      plan->contribs.push_back(c);
      plan->flops += 2.0 * xb.I * (double)ncol * nnew * c.kd0 * c.kd1;
    };
    if (!W) {
      add(nullptr);
    } else {
      for (size_t wi = 0; wi < W->blocks.size(); wi++) {
        const Block& wb = W->blocks[wi];
        if (!W->present.empty() && !W->present[wi]) continue;
        bool match = true;
        for (int k = 0; k < nn; k++) match = match && (wb.c[nk + k] == yb.c[newpos[k]]);
        if (match) add(&wb);
      }
    }
    xb.cnum = (int)plan->contribs.size() - xb.cbeg;
    xb.colstart = col;
    col += ncol;
    plan->blocks.push_back(xb);
  }
  plan->ncols = col;
  plan->bytes = 8.0 * ((double)X.logical_elems() + (double)Y.logical_elems());
  if (W) {
    // groups of blocks with a common source: stable sort by key, consecutive runs become groups
    std::vector<size_t> ord(plan->blocks.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    std::vector<XfBlock> sorted;
    int64_t c0 = 0;
    for (size_t k = 0; k < ord.size(); k++) {
      XfBlock b = plan->blocks[ord[k]];
      int64_t P = 1;
      for (int q = 0; q < MAXP; q++) P *= b.pd[q];
      const bool same = k > 0 && keys[ord[k]] == keys[ord[k - 1]] && plan->groups.back().P == P;
      if (!same) {
        plan->groups.push_back(XfGroup{0, P, (int)sorted.size(), 0, 1, 32});
      }
      plan->groups.back().nb++;
      b.colstart = c0;
      c0 += P;
      sorted.push_back(b);
    }
    plan->blocks.swap(sorted);
    TNL_CHECK(c0 == plan->ncols, "transform groups do not cover the columns");
    int64_t it = 0;
    for (XfGroup& g : plan->groups) {
      g.colstart = it;
      it += (int64_t)g.nb * ((g.P + g.cpw - 1) / g.cpw);
    }
    plan->nitems = it;
  }
  finalize_transform_plan(*plan, W != nullptr);
  return plan;
}

void finalize_transform_plan(TransformPlan& p, bool has_w) {
  p.flats.clear();
  for (XfBlock& b : p.blocks) {
    b.fbeg = (int)p.flats.size();
    for (int c = b.cbeg; c < b.cbeg + b.cnum; c++) {
      const XfContrib& cc = p.contribs[c];
      const int Ka = cc.kd0 * cc.kd1;
      for (int a1 = 0; a1 < cc.kd1; a1++)
        for (int a0 = 0; a0 < cc.kd0; a0++) {
          XfFlat f{};
          f.xoff = cc.xoff + a0 * cc.ks0 + a1 * cc.ks1;
          for (int k = 0; k < MAXP; k++) f.xps[k] = cc.xps[k];
          f.woff = cc.woff + a0 + (int64_t)cc.kd0 * a1;
          f.wst = Ka;
          p.flats.push_back(f);
        }
    }
    b.fnum = (int)p.flats.size() - b.fbeg;
  }
  p.pure_copy = !has_w;
  p.d_groups = p.ctx->upload(p.groups);
  p.d_blocks = p.ctx->upload(p.blocks);
  p.d_flats = p.ctx->upload(p.flats);
}

std::unique_ptr<TransformPlan> plan_scatter(const Tensor& X, Tensor& Y, const SliceMap& slice) {
  auto plan = std::make_unique<TransformPlan>();
  plan->ctx = Y.ctx;
  const int r = X.rank(), pos = slice.ypos;
  TNL_CHECK(r == Y.rank() && pos >= 0 && pos < r && r - 1 <= MAXP, "plan_scatter: bad ranks / slice position");
  int64_t col = 0;
  for (const Block& xb : X.blocks) {
    int yc[MAXR];
    for (int k = 0; k < r; k++) yc[k] = xb.c[k];
    yc[pos] = slice.orig[xb.c[pos]];
    int bi = Y.find(yc);
    TNL_CHECK(bi >= 0, "plan_scatter: block missing in the full tensor");
    const Block& yb = Y.blocks[bi];
    XfBlock b{};
    b.yoff = yb.off + slice.start[xb.c[pos]] * yb.st[pos];
    b.I = xb.d[0];
    b.nd0 = b.nd1 = 1;
    b.yns[0] = b.yns[1] = 0;
    int64_t ncol = 1;
    XfContrib c{};
    c.xoff = xb.off;
    c.kd0 = c.kd1 = 1; c.ks0 = c.ks1 = 0; c.woff = 0;
    for (int k = 0; k < MAXP; k++) { b.pd[k] = 1; b.yps[k] = 0; c.xps[k] = 0; }
    for (int k = 1; k < r; k++) {
      b.pd[k - 1] = xb.d[k];
      b.yps[k - 1] = yb.st[k];
      c.xps[k - 1] = xb.st[k];
      ncol *= xb.d[k];
    }
    b.cbeg = (int)plan->contribs.size();
    b.cnum = 1;
    plan->contribs.push_back(c);
    b.colstart = col;
    col += ncol;
    plan->blocks.push_back(b);
  }
  plan->ncols = col;
  plan->bytes = 16.0 * (double)X.logical_elems();
  finalize_transform_plan(*plan, false);
  return plan;
}

}  // namespace tnl
