// Device StateEnvs{ProjMPO}: environments, H_eff apply, Lanczos.  Reference call sites:
//   position!            src/mps/state_envs.jl:364-367  -> ITensorMPS ProjMPO makeL!/makeR!
//   product / callable   src/mps/state_envs.jl:376-378  -> ProjMPO.product: noprime(v*L*W_j*W_{j+1}*R)
//   eig_solver           src/base/solver.jl:23-43       -> KrylovKit.eigsolve (Lanczos)
//   phi = psi[j]*psi[j+1]  src/mps/update_site.jl:46
// Contraction order follows the reference (L first, then the site operators, then R); in the
// charge-fused layout each big step is a grouped DGEMM over charge sectors and each site-operator
// step is one transform pass.
#include <functional>

#include "env.hpp"
#include "krylov.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <numeric>

namespace tnl {

// ------------------------------------------------------------------------------ tensor helpers
static TensorP mk(Ctx* ctx, std::vector<Index> inds, int nrow, bool cplx = false) {
  return std::make_shared<Tensor>(ctx, std::move(inds), nrow, true, cplx);
}

static Index with_dir(Index ix, int dir) { ix.dir = dir; return ix; }

// tensor whose buffer lives in a persistent workspace slot of the context (zero-filled: pads must be zero)
static TensorP mk_ws(Ctx* ctx, std::vector<Index> inds, int nrow, int slot, bool cplx = false) {
  auto t = std::make_shared<Tensor>(ctx, std::move(inds), nrow, false, cplx);
  t->d = ctx->scratch(slot, (size_t)(t->planes() * t->nelem), true);
  t->owns = false;
  return t;
}

TensorP import_tensor(Ctx* ctx, const HostBlocks& hb, int nrow) {
  TNL_CHECK(hb.rank >= 1 && hb.rank <= MAXR, "import: bad rank");
  auto Y = mk(ctx, hb.inds, nrow, hb.cplx);
  Tensor X(ctx, hb.inds, hb.rank, false);
  X.blocks.clear();
  X.lut.clear();
  X.groups.clear();
  int64_t total = 0;
  for (size_t n = 0; n < hb.coords.size(); n++) {
    Block b{};
    int64_t st = 1;
    for (int k = 0; k < hb.rank; k++) {
      int c = hb.coords[n][k];
      TNL_CHECK(c >= 0 && c < hb.inds[k].nsect(), "import: block coordinate out of range");
      b.c[k] = c; b.d[k] = hb.inds[k].dims[c]; b.st[k] = st; st *= b.d[k];
    }
    b.off = hb.offsets[n];
    b.group = 0;
    TNL_CHECK(Y->find(b.c) >= 0, "import: block violates flux 0 for the given arrows");
    X.lut[Tensor::key(b.c, hb.rank)] = (int)X.blocks.size();
    X.blocks.push_back(b);
    total = std::max(total, b.off + st);
  }
  double* tmp = (double*)ctx->alloc(std::max<int64_t>(total, 1) * sizeof(double));
  std::vector<int> xmap(hb.rank);
  std::iota(xmap.begin(), xmap.end(), 0);
  auto plan = plan_transform(X, *Y, xmap, nullptr, {});
  if (hb.cplx) {
    // host data is interleaved ComplexF64 (NDTensors layout); the device keeps two planes
    std::vector<double> plane((size_t)std::max<int64_t>(total, 1));
    for (int pl = 0; pl < 2; pl++) {
      for (int64_t e = 0; e < total; e++) plane[e] = hb.data[2 * e + pl];
      if (total) CUDA_OK(cudaMemcpyAsync(tmp, plane.data(), total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      run_transform(ctx, *plan, tmp, pl == 0 ? Y->d : Y->im(), nullptr);
      ctx->sync();
    }
  } else {
    if (total) CUDA_OK(cudaMemcpyAsync(tmp, hb.data, total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    run_transform(ctx, *plan, tmp, Y->d, nullptr);
  }
  Y->present.resize(Y->blocks.size());
  for (size_t i = 0; i < Y->blocks.size(); i++) Y->present[i] = X.find(Y->blocks[i].c) >= 0;
  ctx->sync();
  ctx->free(tmp);
  return Y;
}

TensorP relayout(Ctx* ctx, const Tensor& t, int nrow) {
  auto Y = mk(ctx, t.inds, nrow, t.cplx);
  std::vector<int> xmap(t.rank());
  std::iota(xmap.begin(), xmap.end(), 0);
  auto plan = plan_transform(t, *Y, xmap, nullptr, {});
  run_transform_c(ctx, *plan, t, *Y, nullptr);
  ctx->sync();   // the plan's device arrays die with `plan`
  return Y;
}
TensorP to_natural(Ctx* ctx, const Tensor& t) { return relayout(ctx, t, t.rank()); }

// relayout into a persistent workspace slot (transient result)
static TensorP relayout_ws(Ctx* ctx, const Tensor& t, int nrow, int slot) {
  auto Y = std::make_shared<Tensor>(ctx, t.inds, nrow, false, t.cplx);
  Y->d = ctx->scratch(slot, (size_t)(Y->planes() * Y->nelem), true);
  Y->owns = false;
  std::vector<int> xmap(t.rank());
  std::iota(xmap.begin(), xmap.end(), 0);
  auto plan = plan_transform(t, *Y, xmap, nullptr, {});
  run_transform_c(ctx, *plan, t, *Y, nullptr);
  ctx->sync();
  return Y;
}

static TensorP as_nrow(Ctx* ctx, const TensorP& t, int nrow) { return t->nrow == nrow ? t : relayout(ctx, *t, nrow); }

// host permutation of a tiny tensor (MPO site operators)
struct OwnedHost { HostBlocks hb; std::vector<double> data; };
static OwnedHost permute_host(const HostBlocks& h, const std::vector<int>& perm) {
  OwnedHost o;
  o.hb.rank = h.rank;
  for (int k = 0; k < h.rank; k++) o.hb.inds.push_back(h.inds[perm[k]]);
  int64_t off = 0;
  for (size_t n = 0; n < h.coords.size(); n++) {
    std::vector<int> c(h.rank), d(h.rank), dn(h.rank);
    int64_t sz = 1;
    for (int k = 0; k < h.rank; k++) { d[k] = h.inds[k].dims[h.coords[n][k]]; sz *= d[k]; }
    for (int k = 0; k < h.rank; k++) { c[k] = h.coords[n][perm[k]]; dn[k] = d[perm[k]]; }
    o.hb.coords.push_back(c);
    o.hb.offsets.push_back(off);
    o.data.resize(off + sz);
    std::vector<int64_t> sst(h.rank);
    int64_t st = 1;
    for (int k = 0; k < h.rank; k++) { sst[k] = st; st *= d[k]; }
    std::vector<int> idx(h.rank, 0);
    for (int64_t e = 0; e < sz; e++) {
      int64_t src = 0;
      for (int k = 0; k < h.rank; k++) src += idx[k] * sst[perm[k]];
      o.data[off + e] = h.data[h.offsets[n] + src];
      for (int k = 0; k < h.rank; k++) { if (++idx[k] < dn[k]) break; idx[k] = 0; }
    }
    off += sz;
  }
  o.hb.data = o.data.data();
  return o;
}

// -------------------------------------------------------------------------------------- Env
void Env::set_site_op(int site, const HostBlocks& hb) {
  TNL_CHECK(site >= 1 && site <= N, "site out of range");
  TNL_CHECK(hb.rank == 4, "MPO tensor must be (wl, s', s, wr)");
  auto lr = permute_host(hb, {0, 2, 1, 3});     // (wl, s | s', wr)
  auto rl = permute_host(hb, {1, 3, 0, 2});     // (s', wr | wl, s)
  Wlr[site - 1] = import_tensor(ctx, lr.hb, 4);
  Wrl[site - 1] = import_tensor(ctx, rl.hb, 4);
  auto nr = permute_host(hb, {2, 3, 0, 1});     // (s, wr | wl, s')
  Wnr[site - 1] = import_tensor(ctx, nr.hb, 4);
  Wl[site - 1] = hb.inds[0];
  Wr[site - 1] = hb.inds[3];
  lpos = 0; rpos = N + 1; ap.reset(); Ledge.reset(); Redge.reset();
}

// updateH!(sysenv, H; recalcEnv = false) (src/mps/state_envs.jl:181-208): the site operator changes, the cached
// right environments are reused as they are; the MPO links must span the same spaces and the left watermark must be 0
void Env::update_site_op(int site, const HostBlocks& hb) {
  TNL_CHECK(site >= 1 && site <= N && Wlr[site - 1], "updateH!: no operator on this site yet");
  TNL_CHECK(lpos == 0, "updateH!(recalcEnv = false) needs lpos == 0");
  TNL_CHECK(hb.rank == 4 && hb.inds[0].same_space(Wl[site - 1]) && hb.inds[3].same_space(Wr[site - 1]),
            "updateH!(recalcEnv = false): the MPO links must keep their spaces");
  const int lp = lpos, rp = rpos;
  TensorP le = Ledge, re = Redge;
  set_site_op(site, hb);
  lpos = lp; rpos = rp; Ledge = le; Redge = re;
}

Env& Env::term(int k) {
  TNL_CHECK(!parent, "terms are addressed through the top environment");
  TNL_CHECK(k >= 0 && k < 64, "MPO term index out of range");
  while ((int)more.size() < k) more.push_back(std::make_unique<Env>(ctx, N, this));
  return k == 0 ? *this : *more[k - 1];
}

void Env::set_nsite(int n) {
  nsite = n;
  for (auto& m : more) m->nsite = n;
}

void Env::set_state(int site, TensorP a) {
  TNL_CHECK(!parent, "the state belongs to the top environment");
  TNL_CHECK(site >= 1 && site <= N, "site out of range");
  TNL_CHECK(a->rank() == 3, "site tensor must be (l, s, r)");
  A[site - 1] = std::move(a);
  invalidate(site, site);
}

// sites lo..hi of the state changed: environments containing them are stale (watermark semantics of
// src/mps/projcouplingmodel.jl:123-129,212-218)
void Env::invalidate(int lo, int hi) {
  lpos = std::min(lpos, lo - 1);
  rpos = std::max(rpos, hi + 1);
  for (auto& p : pens) {
    p.lpos = std::min(p.lpos, lo - 1);
    p.rpos = std::max(p.rpos, hi + 1);
    p.m.reset();
    p.mloc.reset();
  }
  ap.reset();
  for (auto& m : more) m->invalidate(lo, hi);
}

static void set_one(Ctx* ctx, Tensor& t) {
  TNL_CHECK(t.blocks.size() == 1, "boundary environment must be a single 1x1x1 block");
  double one = 1.0;
  CUDA_OK(cudaMemcpyAsync(t.d + t.blocks[0].off, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
}

void Env::ensure_edges() {
  if (!Ledge) {
    const Index& l = A[0]->inds[0];
    TNL_CHECK(l.dim() == 1 && Wl[0].dim() == 1, "boundary links must have dimension 1");
    Ledge = mk(ctx, {with_dir(l, l.dir), with_dir(Wl[0], -Wl[0].dir), with_dir(l, -l.dir)}, 2);
    set_one(ctx, *Ledge);
  }
  if (!Redge) {
    const Index& r = A[N - 1]->inds[2];
    TNL_CHECK(r.dim() == 1 && Wr[N - 1].dim() == 1, "boundary links must have dimension 1");
    Redge = mk(ctx, {with_dir(r, r.dir), with_dir(Wr[N - 1], -Wr[N - 1].dir), with_dir(r, -r.dir)}, 1);
    set_one(ctx, *Redge);
  }
}

TensorP Env::lproj() { return lpos <= 0 ? Ledge : LR[lpos - 1]; }
TensorP Env::rproj() { return rpos >= N + 1 ? Redge : LR[rpos - 1]; }

static Index trivial_like(const Index& proto, int dir) {
  Index t;
  t.nq = proto.nq;
  t.dir = dir;
  t.dims = {1};
  t.qns = {charge_zero()};
  return t;
}
bool is_trivial_link(const Index& ix) { return ix.nsect() == 1 && ix.dims[0] == 1 && ix.qns[0] == charge_zero(); }

// view of `src` with a trivial (dim 1, charge 0) index inserted at position `at`; shares src's buffer.  With the
// new index in the same (row / column) group as its neighbours the charge-fused layout is byte-identical
// (build_layout enumerates the same combos in the same order), which is checked.
static TensorP view_with_trivial(Ctx* ctx, const Tensor& src, int at, int nrow, const Index& w0, const double* data) {
  std::vector<Index> inds = src.inds;
  inds.insert(inds.begin() + at, w0);
  auto v = std::make_shared<Tensor>(ctx, inds, nrow, false, src.cplx);
  TNL_CHECK(v->nelem == src.nelem && v->blocks.size() == src.blocks.size(), "trivial-link view changed the layout");
  for (size_t i = 0; i < v->blocks.size(); i++)
    TNL_CHECK(v->blocks[i].off == src.blocks[i].off, "trivial-link view changed the block offsets");
  v->d = const_cast<double*>(data);
  v->owns = false;
  return v;
}

// ---- multi-GPU environment update ---------------------------------------------------------------------------------
// Every rank computes the part of the new environment that belongs to its share of one link index (the same
// tnl_shard_range partition the sharded apply uses) and the parts are exchanged with one all-gather: both GEMMs of
// the update shrink by the world size, the result is replicated again (and bitwise identical on every rank).
struct ShardIx { std::vector<SliceMap> sm; std::vector<Index> loc; };
static ShardIx shard_index(Ctx* ctx, const Index& ix, int ypos) {
  ShardIx sh;
  sh.sm.resize(ctx->world);
  sh.loc.resize(ctx->world);
  for (int k = 0; k < ctx->world; k++) {
    Index l = ix;
    l.dims.clear(); l.qns.clear();
    sh.sm[k].ypos = ypos;
    for (int s = 0; s < ix.nsect(); s++) {
      int st, cnt;
      shard_range(ix.dims[s], ctx->world, s, k, &st, &cnt);
      if (cnt > 0) { l.dims.push_back(cnt); l.qns.push_back(ix.qns[s]); sh.sm[k].orig.push_back(s); sh.sm[k].start.push_back(st); }
    }
    sh.loc[k] = l;
  }
  return sh;
}
static bool env_update_shardable(Ctx* ctx, const Tensor* E, const TensorP& Asite, const Index& ix) {
  return ctx->shard_world() > 1 && E && !E->cplx && !Asite->cplx && ix.dim() >= 64 * (int64_t)ctx->world;
}
// `part` = this rank's slice (index `pos` local, any row grouping); returns the full tensor with `nrow` row indices
static TensorP allgather_slices(Ctx* ctx, const Tensor& part, const Index& full_ix, int pos, int nrow, const ShardIx& sh) {
  std::vector<Index> inds = part.inds;
  std::vector<std::shared_ptr<Tensor>> loc(ctx->world);
  int64_t nloc = 0;
  for (int k = 0; k < ctx->world; k++) {
    inds[pos] = with_dir(sh.loc[k], part.inds[pos].dir);
    loc[k] = std::make_shared<Tensor>(ctx, inds, nrow, false);
    nloc = std::max(nloc, loc[k]->nelem);
  }
  nloc = (nloc + 1) & ~int64_t(1);
  double* send = ctx->scratch(Ctx::SLOT_LOCOUT, (size_t)nloc, true);
  double* recv = ctx->scratch(Ctx::SLOT_PACKED, (size_t)nloc * ctx->world, false);
  Tensor& mine = *loc[ctx->rank];
  std::vector<int> id(part.rank());
  std::iota(id.begin(), id.end(), 0);
  auto xr = plan_transform(part, mine, id, nullptr, {});
  run_transform(ctx, *xr, part.d, send, nullptr);
  comm_allgather(ctx, send, recv, nloc);
  inds[pos] = with_dir(full_ix, part.inds[pos].dir);
  auto full = mk(ctx, inds, nrow);
  std::vector<std::unique_ptr<TransformPlan>> keep;
  for (int k = 0; k < ctx->world; k++) {
    SliceMap sm = sh.sm[k];
    sm.ypos = pos;
    keep.push_back(plan_scatter(*loc[k], *full, sm));
    run_transform(ctx, *keep.back(), recv + (int64_t)k * nloc, full->d, nullptr);
  }
  ctx->sync();
  return full;
}

// L_j = L_{j-1} * A_j * W_j * dag(prime(A_j))      (ITensorMPS ProjMPO._makeL!; the same three contractions per id
// in ProjCouplingModel._makeL!, src/mps/projcouplingmodel.jl:123-196).  L == nullptr: the term starts at this
// site (src/mps/projcouplingmodel.jl:151-165, "uncommon tensor" branch with the site operator) -- L is the
// identity on the link and W's left link is trivial, so the first GEMM is skipped.
TensorP Env::step_left(const Tensor* L, const TensorP& Asite, const Tensor& W) {
  TensorP Aq = as_nrow(ctx, Asite, 1);            // [l | s r]
  TensorP As = as_nrow(ctx, Asite, 2);            // [l s | r]
  ap.reset();                                     // the apply workspaces are reused below
  ctx->slot_epoch++;
  // multi-GPU: this rank takes its share of the KET's right link (the column group of the new environment)
  const bool shard = env_update_shardable(ctx, L, Asite, Asite->inds[2]);
  ShardIx sh;
  const Index rfull = Aq->inds[2];
  if (shard) {
    sh = shard_index(ctx, rfull, 2);
    auto Aloc = mk(ctx, {Aq->inds[0], Aq->inds[1], sh.loc[ctx->rank]}, 1);
    auto xa = plan_transform(*Aq, *Aloc, {0, 1, 2}, nullptr, {}, &sh.sm[ctx->rank]);
    run_transform(ctx, *xa, Aq->d, Aloc->d, nullptr);
    ctx->sync();
    Aq = Aloc;
  }
  TensorP X1;
  if (L) {
    X1 = mk_ws(ctx, {L->inds[0], L->inds[1], Aq->inds[1], Aq->inds[2]}, 2, Ctx::SLOT_T1, L->cplx || Aq->cplx);
    auto g1 = plan_gemm(*L, false, *Aq, false, *X1);
    cgemm(ctx, *g1, *L, false, *Aq, false, *X1);
  } else {
    TNL_CHECK(is_trivial_link(W.inds[0]), "a term that starts here must have a trivial left link");
    X1 = view_with_trivial(ctx, *Aq, 1, 2, trivial_like(W.inds[0], -W.inds[0].dir), Aq->d);
  }
  auto Y1 = mk_ws(ctx, {X1->inds[0], W.inds[2], W.inds[3], X1->inds[3]}, 2, Ctx::SLOT_T2, X1->cplx);
  auto x1 = plan_transform(*X1, *Y1, {0, -1, -1, 3}, &W, {1, 2});
  run_transform_c(ctx, *x1, *X1, *Y1, W.d);
  const Index& r = As->inds[2];
  auto Ln = mk_ws(ctx, {with_dir(r, -r.dir), Y1->inds[2], Y1->inds[3]}, 1, Ctx::SLOT_P, Y1->cplx || As->cplx);
  auto g2 = plan_gemm(*As, true, *Y1, false, *Ln, /*dagA=*/true);
  cgemm(ctx, *g2, *As, /*conj: the bra*/true, *Y1, false, *Ln);
  ctx->sync();
  if (shard) return allgather_slices(ctx, *Ln, rfull, 2, 2, sh);
  return relayout(ctx, *Ln, 2);
}

// R_j = dag(prime(A_j)) * R_{j+1} * W_j * A_j  (mirror image; same sum, bra contracted first).  W in the
// (s', wr | wl, s) form.  R == nullptr: the term ends at this site (identity on the link, trivial right link).
TensorP Env::step_right(const Tensor* R, const TensorP& Asite, const Tensor& W) {
  TensorP Aq = as_nrow(ctx, Asite, 1);
  TensorP As = as_nrow(ctx, Asite, 2);
  ap.reset();
  ctx->slot_epoch++;
  // multi-GPU: this rank takes its share of the BRA's left link (the row group of the new environment)
  const bool shard = env_update_shardable(ctx, R, Asite, Asite->inds[0]);
  ShardIx sh;
  const Index lfull = As->inds[0];
  if (shard) {
    sh = shard_index(ctx, lfull, 0);
    auto Aloc = mk(ctx, {sh.loc[ctx->rank], As->inds[1], As->inds[2]}, 2);
    auto xa = plan_transform(*As, *Aloc, {0, 1, 2}, nullptr, {}, &sh.sm[ctx->rank]);
    run_transform(ctx, *xa, As->d, Aloc->d, nullptr);
    ctx->sync();
    As = Aloc;
  }
  const Index &l = As->inds[0], &s = As->inds[1], &r = As->inds[2];
  TensorP Z, conj_keep;
  if (R) {
    Z = mk_ws(ctx, {with_dir(l, -l.dir), with_dir(s, -s.dir), R->inds[1], R->inds[2]}, 2, Ctx::SLOT_T1, As->cplx || R->cplx);
    auto g1 = plan_gemm(*As, false, *R, false, *Z, /*dagA=*/true);
    cgemm(ctx, *g1, *As, /*conj: the bra*/true, *R, false, *Z);
  } else {
    TNL_CHECK(is_trivial_link(W.inds[1]), "a term that ends here must have a trivial right link");
    // dag(A) has the same block coordinates as A; the transform below matches blocks by coordinates.  A complex
    // A needs its conjugate here: a copy with the imaginary plane negated
    if (As->cplx) {
      TensorP Ac = mk(ctx, As->inds, 2, true);
      vec_copy(ctx, Ac->d, As->d, As->nelem);
      vec_scale_to(ctx, Ac->im(), As->im(), As->nelem, -1.0);
      conj_keep = Ac;
      Z = view_with_trivial(ctx, *Ac, 2, 2, trivial_like(W.inds[1], -W.inds[1].dir), Ac->d);
    } else {
      Z = view_with_trivial(ctx, *As, 2, 2, trivial_like(W.inds[1], -W.inds[1].dir), As->d);
    }
  }
  const Index zr = R ? Z->inds[3] : with_dir(r, -r.dir);
  auto Z2 = mk_ws(ctx, {with_dir(l, -l.dir), W.inds[2], W.inds[3], zr}, 2, Ctx::SLOT_T2, Z->cplx);
  auto x1 = plan_transform(*Z, *Z2, {0, -1, -1, 3}, &W, {1, 2});
  run_transform_c(ctx, *x1, *Z, *Z2, W.d);
  auto Rn = mk_ws(ctx, {Z2->inds[0], Z2->inds[1], Aq->inds[0]}, 2, Ctx::SLOT_P, Z2->cplx || Aq->cplx);
  auto g2 = plan_gemm(*Z2, false, *Aq, true, *Rn);
  cgemm(ctx, *g2, *Z2, false, *Aq, false, *Rn);
  ctx->sync();
  if (shard) return allgather_slices(ctx, *Rn, with_dir(lfull, Rn->inds[0].dir), 0, 1, sh);
  return relayout(ctx, *Rn, 1);
}

void Env::makeL(int k) {
  ensure_edges();
  int ll = lpos;
  if (ll >= k) { lpos = k; return; }
  ll = std::max(ll, 0);
  TensorP L = lproj();
  while (ll < k) {
    L = step_left(L.get(), A[ll], *Wlr[ll]);
    LR[ll] = L;
    ll++;
  }
  lpos = k;
}

void Env::makeR(int k) {
  ensure_edges();
  int rl = rpos;
  if (rl <= k) { rpos = k; return; }
  rl = std::min(rl, N + 1);
  TensorP R = rproj();
  while (rl > k) {
    int j = rl - 2;
    R = step_right(R.get(), A[j], *Wrl[j]);
    LR[j] = R;
    rl--;
  }
  rpos = k;
}

void Env::position(int pos) {
  int lp = lpos, rp = rpos;
  if (cm) {
    cm_makeL(pos - 1);
    cm_makeR(pos + nsite);
  } else {
    makeL(pos - 1);
    makeR(pos + nsite);
  }
  if (lp != lpos || rp != rpos) ap.reset();
  for (auto& p : pens) position_penalty(p, pos);
  for (auto& m : more) { m->nsite = nsite; m->position(pos); }
}

// ------------------------------------------------------------------------ excited-state penalty
// ProjMPS2 (src/mps/projmps2.jl:77-124): L_j = L_{j-1} * psi[j] * dag(prime(M[j], "Link")), mirror for R.
void Env::add_penalty(const std::vector<TensorP>& M, double w) {
  TNL_CHECK((int)M.size() == N, "penalised MPS has the wrong length");
  TNL_CHECK(w > 0.0, "`weight` parameter should be > 0.0");
  Penalty p;
  p.M = M;
  p.LR.resize(N);
  p.rpos = N + 1;
  pens.push_back(std::move(p));
  weight = w;
  ap.reset();
}

void Env::position_penalty(Penalty& p, int pos) {
  if (p.dead) return;
  // ComplexF64 states / penalised states: every contraction below is the planar complex GEMM (dag = arrows + conj)
  if (!p.Ledge) {
    const Index &a0 = A[0]->inds[0], &m0 = p.M[0]->inds[0], &aN = A[N - 1]->inds[2], &mN = p.M[N - 1]->inds[2];
    p.Ledge = mk(ctx, {with_dir(a0, -a0.dir), with_dir(m0, m0.dir)}, 1);
    p.Redge = mk(ctx, {with_dir(aN, -aN.dir), with_dir(mN, mN.dir)}, 1);
    if (p.Ledge->blocks.size() != 1 || p.Redge->blocks.size() != 1) { p.dead = true; return; }
    set_one(ctx, *p.Ledge);
    set_one(ctx, *p.Redge);
  }
  const int kl = pos - 1, kr = pos + nsite;
  // left overlap environments
  if (p.lpos >= kl) {
    p.lpos = kl;
  } else {
    int ll = std::max(p.lpos, 0);
    TensorP L = ll <= 0 ? p.Ledge : p.LR[ll - 1];
    while (ll < kl) {
      TensorP Aq = as_nrow(ctx, A[ll], 1);
      TensorP Ms = as_nrow(ctx, p.M[ll], 2);
      auto X = mk(ctx, {L->inds[1], Aq->inds[1], Aq->inds[2]}, 1, L->cplx || Aq->cplx);   // (lM, s, r)
      auto g1 = plan_gemm(*L, true, *Aq, false, *X);
      cgemm(ctx, *g1, *L, false, *Aq, false, *X);
      ctx->sync();
      TensorP X2 = relayout(ctx, *X, 2);
      const Index& rm = Ms->inds[2];
      auto Ln = mk(ctx, {X->inds[2], with_dir(rm, -rm.dir)}, 1, X2->cplx || Ms->cplx);    // (r, rM)
      auto g2 = plan_gemm(*X2, true, *Ms, false, *Ln, false, /*dagB=*/true);
      cgemm(ctx, *g2, *X2, false, *Ms, true, *Ln);
      ctx->sync();
      p.LR[ll] = Ln;
      L = Ln;
      ll++;
    }
    p.lpos = kl;
  }
  if (p.rpos <= kr) {
    p.rpos = kr;
  } else {
    int rl = std::min(p.rpos, N + 1);
    TensorP R = rl >= N + 1 ? p.Redge : p.LR[rl - 1];
    while (rl > kr) {
      const int j = rl - 2;
      TensorP As = as_nrow(ctx, A[j], 2);
      TensorP Mq = as_nrow(ctx, p.M[j], 1);
      auto Y = mk(ctx, {As->inds[0], As->inds[1], R->inds[1]}, 2, As->cplx || R->cplx);   // (l, s, rM)
      auto g1 = plan_gemm(*As, false, *R, false, *Y);
      cgemm(ctx, *g1, *As, false, *R, false, *Y);
      ctx->sync();
      TensorP Y1 = relayout(ctx, *Y, 1);
      const Index& lm = Mq->inds[0];
      auto Rn = mk(ctx, {Y->inds[0], with_dir(lm, -lm.dir)}, 1, Y1->cplx || Mq->cplx);    // (l, lM)
      auto g2 = plan_gemm(*Y1, false, *Mq, true, *Rn, false, /*dagB=*/true);
      cgemm(ctx, *g2, *Y1, false, *Mq, true, *Rn);
      ctx->sync();
      p.LR[j] = Rn;
      R = Rn;
      rl--;
    }
    p.rpos = kr;
  }
  p.m.reset();
  p.mloc.reset();
}

// |m> = dag(proj_mps) = dag(L) * M_j * M_{j+1} * dag(R)   (src/mps/projmps2.jl:182-196)
void Env::build_penalty_vector(Penalty& p, const Tensor& proto) {
  if (p.dead || p.m) return;
  TensorP L = p.lpos <= 0 ? p.Ledge : p.LR[p.lpos - 1];
  TensorP R = p.rpos >= N + 1 ? p.Redge : p.LR[p.rpos - 1];
  TNL_CHECK(L && R, "penalty environments not positioned");
  TensorP cur;           // running tensor with indices (l, sites..., link_M) , bipartition before the last index
  const int first = p.lpos + 1;
  if (nsite == 0) {
    auto m = mk(ctx, {with_dir(L->inds[0], -L->inds[0].dir), with_dir(R->inds[0], -R->inds[0].dir)}, 1, proto.cplx);
    auto g = plan_gemm(*L, false, *R, true, *m, true, true);
    if (proto.cplx) cgemm(ctx, *g, L->d, L->cplx ? L->im() : nullptr, true, R->d, R->cplx ? R->im() : nullptr, true, m->d, m->im());
    else run_gemm(ctx, *g, L->d, R->d, m->d);
    ctx->sync();
    p.m = m;
  } else {
    TensorP M1 = as_nrow(ctx, p.M[first - 1], 1);
    auto P1 = mk(ctx, {with_dir(L->inds[0], -L->inds[0].dir), M1->inds[1], M1->inds[2]}, 1, L->cplx || M1->cplx);   // (l, s1, mM)
    auto g1 = plan_gemm(*L, false, *M1, false, *P1, /*dagA=*/true);
    cgemm(ctx, *g1, *L, true, *M1, false, *P1);
    ctx->sync();
    cur = P1;
    if (nsite == 2) {
      TensorP P1s = relayout(ctx, *P1, 2);
      TensorP M2 = as_nrow(ctx, p.M[first], 1);
      auto P2 = mk(ctx, {P1->inds[0], P1->inds[1], M2->inds[1], M2->inds[2]}, 2, P1s->cplx || M2->cplx);   // (l, s1, s2, rM)
      auto g2 = plan_gemm(*P1s, false, *M2, false, *P2);
      cgemm(ctx, *g2, *P1s, false, *M2, false, *P2);
      ctx->sync();
      cur = P2;
    }
    const int r = cur->rank();
    TensorP C3 = relayout(ctx, *cur, r - 1);
    std::vector<Index> mi(cur->inds.begin(), cur->inds.end() - 1);
    mi.push_back(with_dir(R->inds[0], -R->inds[0].dir));
    // a real |m> under a complex Krylov vector is stored complex (zero imaginary plane): one layout for cdot / caxpy
    auto m = mk(ctx, mi, r - 1, proto.cplx);
    auto g3 = plan_gemm(*C3, false, *R, true, *m, false, /*dagB=*/true);
    if (proto.cplx) cgemm(ctx, *g3, C3->d, C3->cplx ? C3->im() : nullptr, false, R->d, R->cplx ? R->im() : nullptr, true, m->d, m->im());
    else {
      TNL_CHECK(!C3->cplx && !R->cplx, "complex penalised state under a real Krylov vector: promote the state first");
      run_gemm(ctx, *g3, C3->d, R->d, m->d);
    }
    ctx->sync();
    p.m = relayout(ctx, *m, 1);
  }
  TNL_CHECK(p.m->nelem == proto.nelem, "penalty projector layout differs from the Krylov vector layout");
}

TensorP Env::make_phi(int pos) {
  TNL_CHECK(pos >= 1 && pos < N, "bond out of range");
  TensorP A1 = as_nrow(ctx, A[pos - 1], 2);
  TensorP A2 = as_nrow(ctx, A[pos], 1);
  ap.reset();
  ctx->slot_epoch++;
  auto S = mk_ws(ctx, {A1->inds[0], A1->inds[1], A2->inds[1], A2->inds[2]}, 2, Ctx::SLOT_T3, A1->cplx || A2->cplx);
  auto g = plan_gemm(*A1, false, *A2, false, *S);
  cgemm(ctx, *g, *A1, false, *A2, false, *S);
  ctx->sync();
  return relayout(ctx, *S, 1);
}

// ---------------------------------------------------------------------------------- H_eff apply
struct Env::ApplyPlan {
  int nsite;
  bool sharded = false;              // multi-GPU: this rank handles a slice of the right link r
  TensorP vloc;                      // structure of a Krylov vector restricted to the local r range (no data)
  std::unique_ptr<TransformPlan> xs; // full Krylov layout -> local slice
  int64_t nloc = 0;                  // padded length of a local Krylov vector (same on every rank)
  std::vector<std::unique_ptr<TransformPlan>> pack;    // P -> packed + k*nloc : the slice of rank k
  std::vector<std::unique_ptr<TransformPlan>> unpack;  // packed + k*nloc -> full Krylov layout
  double* packed = nullptr;          // world * nloc doubles
  double* loc_in = nullptr;          // scratch local vectors for the full-vector entry points
  double* loc_out = nullptr;
  std::vector<SliceMap> pack_maps;
  std::vector<TensorP> qloc;
  Ctx* ctx = nullptr;
  // fused `T3 R^T` -> reduce-scatter over peer memory: scatter tables of the GEMM epilogue (device) and their owner
  ScatterProb* d_scatter = nullptr;
  std::vector<void*> dev_allocs;
  ~ApplyPlan() { if (ctx) for (void* q : dev_allocs) ctx->free(q); }

  TensorP L, R, W1, W2;
  TensorP T1, T2, T3, P;
  std::unique_ptr<GemmPlan> g1, g4;
  std::unique_ptr<TransformPlan> x2, x3, x5;
  int64_t nelem;
  double flops;
  uint64_t epoch = 0;                // ctx->slot_epoch when the plan was built (workspace pointers valid)
  bool cplx = false;                 // planar complex vectors: plane distance = nelem
};

void Env::build_apply_plan(const Tensor& vfull) {
  const int first = lpos + 1;      // 1-based first site of the range
  TNL_CHECK(rpos - lpos == nsite + 1, "environments are not positioned for this nsite");
  ap = make_plan(vfull, lproj(), nsite >= 1 ? Wlr[first - 1] : nullptr, nsite == 2 ? Wlr[first] : nullptr, rproj(), true);
}

// Plan of one H_eff term  v -> L * W1 * W2 * R  (any of L / R may be absent: a CouplingModel term that starts or
// ends inside the site range, src/mps/projcouplingmodel.jl:315-356 -- the corresponding GEMM is skipped and the
// operand is a trivial-link VIEW of its neighbour, no flops and no copy).
std::shared_ptr<Env::ApplyPlan> Env::make_plan(const Tensor& vfull, TensorP Lp, TensorP W1p, TensorP W2p, TensorP Rp,
                                               bool allow_shard) {
  struct Timer {
    Ctx* c; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~Timer() { c->cnt.host_plan_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
  } timer{ctx};
  const Tensor& v = vfull;
  auto p = std::make_shared<ApplyPlan>();
  p->nsite = nsite;
  p->L = Lp;
  p->R = Rp;
  const bool cp = v.cplx;
  p->cplx = cp;
  TNL_CHECK(cp || !((Lp && Lp->cplx) || (Rp && Rp->cplx)), "complex environments need a complex vector (promote it first)");
  TNL_CHECK(p->L || p->R || nsite > 0, "a zero-site term needs an environment");
  TNL_CHECK(v.nrow == 1, "Krylov vectors use the [l | rest] layout");
  // link carried by the running tensor before the first / after the last site operator when L / R is absent
  auto win = [&](const TensorP& W) { return trivial_like(W->inds[0], -W->inds[0].dir); };
  if (nsite == 2) {
    TNL_CHECK(v.rank() == 4, "two-site apply needs a (l,s1,s2,r) vector");
    p->W1 = W1p;
    p->W2 = W2p;
    const Tensor* vin = &v;
    // bonds near the chain ends (right link smaller than 64 per rank) run replicated: every rank computes the
    // identical full result, no collective
    // (complex vectors shard plane by plane; penalised complex states stay replicated)
    if (ctx->shard_world() > 1 && allow_shard && p->L && p->R && !(cp && !pens.empty()) && v.inds[3].dim() >= 64 * (int64_t)ctx->world) {
      // Sharded apply: every rank owns a contiguous share of each sector of the right
      // link r.  It contracts L with its slice of v, carries the slice through the site operators and contracts
      // with its slice of R over (w, r_local): a partial H_eff v of full size, summed by one NCCL all-reduce.
      const Index& r = v.inds[3];
      const Tensor& Rf = *p->R;
      p->sharded = true;
      p->ctx = ctx;
      Tensor Qfull(ctx, v.inds, 1, false);
      std::vector<SliceMap> sms(ctx->world);
      std::vector<TensorP> qloc(ctx->world);
      for (int k = 0; k < ctx->world; k++) {
        Index rl = r;
        rl.dims.clear(); rl.qns.clear();
        sms[k].ypos = 3;
        for (int s = 0; s < r.nsect(); s++) {
          int st, cnt;
          shard_range(r.dims[s], ctx->world, s, k, &st, &cnt);
          if (cnt > 0) { rl.dims.push_back(cnt); rl.qns.push_back(r.qns[s]); sms[k].orig.push_back(s); sms[k].start.push_back(st); }
        }
        qloc[k] = std::make_shared<Tensor>(ctx, std::vector<Index>{v.inds[0], v.inds[1], v.inds[2], rl}, 1, false, cp);
        p->nloc = std::max(p->nloc, qloc[k]->nelem);
      }
      const SliceMap& sm = sms[ctx->rank];
      p->vloc = qloc[ctx->rank];
      p->xs = plan_transform(v, *p->vloc, {0, 1, 2, 3}, nullptr, {}, &sm);
      const size_t npl = cp ? 2 : 1;                               // planes of a (planar) complex vector
      p->packed = ctx->scratch(Ctx::SLOT_PACKED, (size_t)p->nloc * ctx->world * npl, true);
      p->loc_in = ctx->scratch(Ctx::SLOT_LOCIN, (size_t)p->nloc * npl, true);
      p->loc_out = ctx->scratch(Ctx::SLOT_LOCOUT, (size_t)p->nloc * npl, true);
      for (int k = 0; k < ctx->world; k++) p->unpack.push_back(plan_scatter(*qloc[k], Qfull, sms[k]));
      // pack plans need P: built below once P exists (same slice maps on P's last index)
      p->pack_maps = sms;
      p->qloc = qloc;
      Index rlR = p->vloc->inds[3];
      rlR.dir = Rf.inds[2].dir;
      SliceMap sr = sm;
      sr.ypos = 2;
      auto Rloc = mk(ctx, {Rf.inds[0], Rf.inds[1], rlR}, 1, Rf.cplx);
      auto xr = plan_transform(Rf, *Rloc, {0, 1, 2}, nullptr, {}, &sr);
      run_transform_c(ctx, *xr, Rf, *Rloc, nullptr);
      ctx->sync();
      p->R = Rloc;
      vin = p->vloc.get();
    }
    const Tensor& v = *vin;     // from here on: the (possibly sliced) input structure
    const Tensor &W1 = *p->W1, &W2 = *p->W2;
    const Index lout = p->L ? p->L->inds[0] : v.inds[0];
    if (p->L) {
      const Tensor& L = *p->L;
      p->T1 = mk_ws(ctx, {L.inds[0], L.inds[1], v.inds[1], v.inds[2], v.inds[3]}, 2, Ctx::SLOT_T1, cp);
      p->g1 = plan_gemm(L, false, v, false, *p->T1);
    } else {
      TNL_CHECK(is_trivial_link(W1.inds[0]), "term without a left environment must start with a trivial link");
      p->T1 = view_with_trivial(ctx, v, 1, 2, win(p->W1), nullptr);          // data = the input vector itself
    }
    p->T2 = mk_ws(ctx, {lout, W1.inds[2], W1.inds[3], v.inds[2], v.inds[3]}, 5, Ctx::SLOT_T2, cp);
    p->x2 = plan_transform(*p->T1, *p->T2, {0, -1, -1, 3, 4}, &W1, {1, 2});
    p->T3 = mk_ws(ctx, {lout, W1.inds[2], W2.inds[2], W2.inds[3], v.inds[3]}, 3, Ctx::SLOT_T3, cp);
    p->x3 = plan_transform(*p->T2, *p->T3, {0, 1, -1, -1, 4}, &W2, {2, 3});
    if (p->R) {
      const Tensor& R = *p->R;
      p->P = mk_ws(ctx, {lout, W1.inds[2], W2.inds[2], R.inds[0]}, 3, Ctx::SLOT_P, cp);
      p->g4 = plan_gemm(*p->T3, false, R, true, *p->P);
    } else {
      TNL_CHECK(is_trivial_link(W2.inds[3]), "term without a right environment must end with a trivial link");
      p->P = std::make_shared<Tensor>(ctx, std::vector<Index>{lout, W1.inds[2], W2.inds[2], v.inds[3]}, 3, false, cp);
      TNL_CHECK(p->P->nelem == p->T3->nelem, "trivial-link view changed the layout");
      p->P->d = p->T3->d;
      p->P->owns = false;
    }
    Tensor Q(ctx, p->P->inds, 1, false);
    TNL_CHECK(Q.nelem == vfull.nelem, "H_eff output layout differs from the input layout (index mismatch)");
    p->x5 = plan_transform(*p->P, Q, {0, 1, 2, 3}, nullptr, {});
    if (p->sharded) {
      for (int k = 0; k < ctx->world; k++)
        p->pack.push_back(plan_transform(*p->P, *p->qloc[k], {0, 1, 2, 3}, nullptr, {}, &p->pack_maps[k]));
      build_scatter_tables(*p);
    }
    p->flops = (p->g1 ? p->g1->flops : 0.0) + p->x2->flops + p->x3->flops + (p->g4 ? p->g4->flops : 0.0);
  } else if (nsite == 1) {
    TNL_CHECK(v.rank() == 3, "one-site apply needs a (l,s,r) vector");
    p->W1 = W1p;
    const Tensor& W1 = *p->W1;
    const Index lout = p->L ? p->L->inds[0] : v.inds[0];
    if (p->L) {
      const Tensor& L = *p->L;
      p->T1 = mk_ws(ctx, {L.inds[0], L.inds[1], v.inds[1], v.inds[2]}, 2, Ctx::SLOT_T1, cp);
      p->g1 = plan_gemm(L, false, v, false, *p->T1);
    } else {
      TNL_CHECK(is_trivial_link(W1.inds[0]), "term without a left environment must start with a trivial link");
      p->T1 = view_with_trivial(ctx, v, 1, 2, win(p->W1), nullptr);
    }
    p->T3 = mk_ws(ctx, {lout, W1.inds[2], W1.inds[3], v.inds[2]}, 2, Ctx::SLOT_T3, cp);
    p->x2 = plan_transform(*p->T1, *p->T3, {0, -1, -1, 3}, &W1, {1, 2});
    if (p->R) {
      const Tensor& R = *p->R;
      p->P = mk_ws(ctx, {lout, W1.inds[2], R.inds[0]}, 2, Ctx::SLOT_P, cp);
      p->g4 = plan_gemm(*p->T3, false, R, true, *p->P);
    } else {
      TNL_CHECK(is_trivial_link(W1.inds[3]), "term without a right environment must end with a trivial link");
      p->P = std::make_shared<Tensor>(ctx, std::vector<Index>{lout, W1.inds[2], v.inds[2]}, 2, false, cp);
      TNL_CHECK(p->P->nelem == p->T3->nelem, "trivial-link view changed the layout");
      p->P->d = p->T3->d;
      p->P->owns = false;
    }
    Tensor Q(ctx, p->P->inds, 1, false);
    TNL_CHECK(Q.nelem == v.nelem, "H_eff output layout differs from the input layout (index mismatch)");
    p->x5 = plan_transform(*p->P, Q, {0, 1, 2}, nullptr, {});
    p->flops = (p->g1 ? p->g1->flops : 0.0) + p->x2->flops + (p->g4 ? p->g4->flops : 0.0);
  } else if (nsite == 0) {
    TNL_CHECK(v.rank() == 2, "zero-site apply needs a (l,r) bond matrix");
    if (p->L && p->R) {
      const Tensor &L = *p->L, &R = *p->R;
      p->T1 = mk_ws(ctx, {L.inds[0], L.inds[1], v.inds[1]}, 2, Ctx::SLOT_T1, cp);            // [(l' w) | r]
      p->g1 = plan_gemm(L, false, v, false, *p->T1);
      p->T3 = mk_ws(ctx, p->T1->inds, 1, Ctx::SLOT_T3, cp);                                  // [l' | (w r)]
      p->x2 = plan_transform(*p->T1, *p->T3, {0, 1, 2}, nullptr, {});
      p->P = mk_ws(ctx, {L.inds[0], R.inds[0]}, 1, Ctx::SLOT_P, cp);
      p->g4 = plan_gemm(*p->T3, false, R, true, *p->P);
      TNL_CHECK(p->P->nelem == v.nelem, "H_eff output layout differs from the input layout (index mismatch)");
      p->flops = p->g1->flops + p->g4->flops;
    } else if (p->L) {
      // out[l' | r] = L[(l' w0) | l] v[l | r]   (local left environment: the link is trivial)
      const Tensor& L = *p->L;
      TNL_CHECK(is_trivial_link(L.inds[1]), "left-only zero-site term must carry a trivial link");
      p->P = std::make_shared<Tensor>(ctx, std::vector<Index>{L.inds[0], L.inds[1], v.inds[1]}, 2, false, cp);
      TNL_CHECK(p->P->nelem == v.nelem, "H_eff output layout differs from the input layout (index mismatch)");
      p->g1 = plan_gemm(L, false, v, false, *p->P);
      p->flops = p->g1->flops;
    } else {
      // out[l | r'] = v[l | (w0 r)] R[r' | (w0 r)]^T
      const Tensor& R = *p->R;
      TNL_CHECK(is_trivial_link(R.inds[1]), "right-only zero-site term must carry a trivial link");
      p->T3 = view_with_trivial(ctx, v, 1, 1, with_dir(R.inds[1], -R.inds[1].dir), nullptr);
      p->P = std::make_shared<Tensor>(ctx, std::vector<Index>{v.inds[0], R.inds[0]}, 1, false, cp);
      TNL_CHECK(p->P->nelem == v.nelem, "H_eff output layout differs from the input layout (index mismatch)");
      p->g4 = plan_gemm(*p->T3, false, R, true, *p->P);
      p->flops = p->g4->flops;
    }
  } else {
    throw Error(2, "nsite must be 0, 1 or 2");
  }
  if (cp) {
    // planar complex arithmetic: a (complex environment) x (complex tensor) product is four real GEMM launches (8mnk),
    // a real environment two (4mnk); the site-operator passes run once per plane
    const double fl = p->L && p->L->cplx ? 4.0 : 2.0, fr = p->R && p->R->cplx ? 4.0 : 2.0;
    const double xf = p->flops - (p->g1 ? p->g1->flops : 0.0) - (p->g4 ? p->g4->flops : 0.0);
    p->flops = fl * (p->g1 ? p->g1->flops : 0.0) + fr * (p->g4 ? p->g4->flops : 0.0) + 2.0 * xf;
  }
  p->nelem = vfull.nelem;
  p->epoch = ctx->slot_epoch;
  return p;
}

// Scatter tables of the fused reduce-scatter: for every charge-sector problem of `T3 R^T` (a group of
// P[(l' s1' s2') | r']) where each output element lives in the local Krylov layout [l | (s1 s2 r_k)] of the rank k
// that owns its column.  Row side: the row block (l' sector, s1', s2') and the offset inside the destination block;
// column side: owner rank and column inside its share.  Falls back to pack + ncclReduceScatter when the staging
// area cannot be set up or the plan has tiles the TMA kernel does not serve.
void Env::build_scatter_tables(ApplyPlan& p) {
  if (!p.g4 || !ctx->use_tma || p.g4->probs.empty() || p.cplx) return;   // complex: four accumulating launches, NCCL path
  if (!comm_stage_ensure(ctx, (size_t)p.nloc)) return;
  gemm_plan_force_tma(*p.g4);
  const Tensor& P = *p.P;
  const int W = ctx->world;
  std::vector<int2> rowinfo, colinfo;
  std::vector<int64_t> Tt;
  std::vector<int> cstride;
  struct Offs { size_t row, col, t, cs; int nseg; };
  std::vector<Offs> offs;
  for (const GemmProblem& gp : p.g4->probs) {
    int gi = -1;
    for (size_t g = 0; g < P.groups.size(); g++) if (P.groups[g].base == gp.c) gi = (int)g;
    TNL_CHECK(gi >= 0, "scatter tables: GEMM problem without an output group");
    const Group& G = P.groups[gi];
    TNL_CHECK(G.R == gp.M && G.C == gp.N, "scatter tables: group shape");
    const int nrb = (int)G.rows.size();
    const int nseg = (int)G.cols.size() * W;
    TNL_CHECK(nseg < 65536 && W < 32768, "scatter tables: too many column segments");
    Offs o{rowinfo.size(), colinfo.size(), Tt.size(), cstride.size(), nseg};
    rowinfo.resize(o.row + (size_t)G.R);
    colinfo.resize(o.col + (size_t)G.C);
    Tt.resize(o.t + (size_t)nrb * nseg, -1);
    cstride.resize(o.cs + (size_t)nrb, 0);
    // destination blocks: (row combo, column sector, rank)
    std::vector<int64_t> st1(nrb, -1), st2(nrb, -1);
    for (int rb = 0; rb < nrb; rb++) {
      const Combo& rc = G.rows[rb];
      for (size_t cc = 0; cc < G.cols.size(); cc++) {
        const int bsec = G.cols[cc].c[0];
        for (int k = 0; k < W; k++) {
          const SliceMap& sm = p.pack_maps[k];
          int bloc = -1;
          for (size_t q = 0; q < sm.orig.size(); q++) if (sm.orig[q] == bsec) bloc = (int)q;
          if (bloc < 0) continue;                  // rank k holds nothing of this sector
          int co[MAXR] = {rc.c[0], rc.c[1], rc.c[2], bloc};
          const int bi = p.qloc[k]->find(co);
          TNL_CHECK(bi >= 0, "scatter tables: destination block missing");
          const Block& db = p.qloc[k]->blocks[bi];
          Tt[o.t + (size_t)rb * nseg + cc * W + k] = db.off;
          TNL_CHECK(db.st[0] == 1, "scatter tables: destination rows must be contiguous");
          if (st1[rb] < 0) { st1[rb] = db.st[1]; st2[rb] = db.st[2]; cstride[o.cs + rb] = (int)db.st[3]; }
          TNL_CHECK(st1[rb] == db.st[1] && st2[rb] == db.st[2] && cstride[o.cs + rb] == (int)db.st[3],
                    "scatter tables: destination strides differ between ranks");
        }
      }
      for (int64_t e = 0; e < rc.dim; e++) {
        const int64_t i0 = e % rc.d[0], i1 = (e / rc.d[0]) % rc.d[1], i2 = e / ((int64_t)rc.d[0] * rc.d[1]);
        const int64_t roff = i0 + i1 * std::max<int64_t>(st1[rb], 0) + i2 * std::max<int64_t>(st2[rb], 0);
        TNL_CHECK(roff < INT32_MAX, "scatter tables: row offset overflow");
        rowinfo[o.row + (size_t)(rc.off + e)] = make_int2(rb, (int)roff);
      }
    }
    for (size_t cc = 0; cc < G.cols.size(); cc++) {
      const Combo& c = G.cols[cc];
      const int bsec = c.c[0];
      for (int k = 0; k < W; k++) {
        int st, cnt;
        shard_range((int)c.dim, W, bsec, k, &st, &cnt);
        for (int j = 0; j < cnt; j++) colinfo[o.col + (size_t)(c.off + st + j)] = make_int2((int)(cc * W + k) | (k << 16), j);
      }
    }
    offs.push_back(o);
  }
  p.ctx = ctx;
  int2* d_row = ctx->upload(rowinfo);
  int2* d_col = ctx->upload(colinfo);
  int64_t* d_T = ctx->upload(Tt);
  int* d_cs = ctx->upload(cstride);
  std::vector<ScatterProb> sp(offs.size());
  for (size_t i = 0; i < offs.size(); i++)
    sp[i] = ScatterProb{d_row + offs[i].row, d_col + offs[i].col, d_T + offs[i].t, d_cs + offs[i].cs, offs[i].nseg, 0};
  p.d_scatter = ctx->upload(sp);
  p.dev_allocs = {d_row, d_col, d_T, d_cs, p.d_scatter};
}

double Env::apply_flops() const {
  if (cm) return cm_apply_flops();
  double f = ap ? ap->flops : 0.0;
  for (auto& m : more) f += m->apply_flops();
  return f;
}

void Env::apply(const Tensor& v, Tensor& out) {
  TNL_CHECK(out.nelem == v.nelem && out.nrow == 1 && out.cplx == v.cplx, "output vector layout mismatch");
  apply_ptr(v, v.d, out.d);
}

// Sharded core: local slice in, local slice of H_eff v out (sum over ranks by reduce-scatter).
void Env::apply_local(const double* vloc, double* outloc) {
  ApplyPlan& p = *ap;
  if (p.cplx) {
    // planar complex vectors: local planes at distance nloc; every contraction is 2-4 launches of the real kernels,
    // each plane is packed and reduce-scattered on its own
    const int64_t nl = p.nloc, W = ctx->world;
    auto imc = [&](const TensorP& t) -> const double* { return t->cplx ? t->im() : nullptr; };
    cgemm(ctx, *p.g1, p.L->d, imc(p.L), false, vloc, vloc + nl, false, p.T1->d, p.T1->im());
    run_transform_c(ctx, *p.x2, *p.T1, *p.T2, p.W1->d);
    run_transform_c(ctx, *p.x3, *p.T2, *p.T3, p.W2->d);
    cgemm(ctx, *p.g4, p.T3->d, p.T3->im(), false, p.R->d, imc(p.R), false, p.P->d, p.P->im());
    for (int pl = 0; pl < 2; pl++) {
      const double* src = pl ? p.P->im() : p.P->d;
      double* pk = p.packed + pl * W * nl;
      for (int k = 0; k < W; k++) run_transform(ctx, *p.pack[k], src, pk + (int64_t)k * nl, nullptr);
      comm_reduce_scatter_sum(ctx, pk, outloc + pl * nl, nl);
    }
    if (!more.empty()) {
      double* tmp = ctx->vec_acquire((size_t)(2 * nl));
      for (auto& m : more) {
        m->apply_local(vloc, tmp);
        vec_axpy(ctx, outloc, tmp, 2 * nl, 1.0);
      }
      ctx->vec_release(tmp, 0);
    }
    if (!parent) ctx->cnt.apply_count += 1;
    return;
  }
  run_gemm(ctx, *p.g1, p.L->d, vloc, p.T1->d);
  run_transform(ctx, *p.x2, p.T1->d, p.T2->d, p.W1->d);
  run_transform(ctx, *p.x3, p.T2->d, p.T3->d, p.W2->d);
  if (p.d_scatter) {
    // one kernel does the GEMM and the exchange: tiles go straight into the owners' staging slots over NVLink
    run_gemm_reduce_scatter(ctx, *p.g4, p.T3->d, p.R->d, p.d_scatter, outloc, p.nloc);
  } else {
    run_gemm(ctx, *p.g4, p.T3->d, p.R->d, p.P->d);
    for (int k = 0; k < ctx->world; k++) run_transform(ctx, *p.pack[k], p.P->d, p.packed + (int64_t)k * p.nloc, nullptr);
    comm_reduce_scatter_sum(ctx, p.packed, outloc, p.nloc);
  }
  if (!more.empty()) {
    double* tmp = ctx->vec_acquire((size_t)p.nloc);
    for (auto& m : more) {
      m->apply_local(vloc, tmp);
      vec_axpy(ctx, outloc, tmp, p.nloc, 1.0);
    }
    ctx->vec_release(tmp, 0);
  }
  for (auto& pen : pens) {
    if (pen.dead) continue;
    if (!pen.mloc) {
      pen.mloc = std::make_shared<Tensor>(ctx, p.vloc->inds, 1, false);
      pen.mloc->d = (double*)ctx->alloc((size_t)p.nloc * sizeof(double));
      CUDA_OK(cudaMemsetAsync(pen.mloc->d, 0, (size_t)p.nloc * sizeof(double), ctx->stream));
      run_transform(ctx, *p.xs, pen.m->d, pen.mloc->d, nullptr);
    }
    vec_dot(ctx, pen.mloc->d, vloc, p.nloc, 200);
    comm_allreduce_sum(ctx, ctx->d_scalars + 200, 1);
    vec_axpy_dev(ctx, outloc, pen.mloc->d, p.nloc, 200, weight);
  }
  if (!parent) ctx->cnt.apply_count += 1;
}

void Env::ensure_plan(const Tensor& proto) {
  if (cm) { cm_ensure_plans(proto); return; }
  // the terms of an MPO sum share the context's workspace slots: a plan is stale when a slot was reallocated
  // (or reused by another phase) after it was built; two passes reach a fixed point because capacities only grow
  for (int pass = 0; pass < 3; pass++) {
    bool rebuilt = false;
    if (!ap || ap->nsite != nsite || ap->nelem != proto.nelem || ap->cplx != proto.cplx || ap->epoch != ctx->slot_epoch) { build_apply_plan(proto); rebuilt = true; }
    for (auto& m : more) {
      m->nsite = nsite;
      if (!m->ap || m->ap->nsite != nsite || m->ap->nelem != proto.nelem || m->ap->cplx != proto.cplx || m->ap->epoch != ctx->slot_epoch) { m->build_apply_plan(proto); rebuilt = true; }
    }
    if (!rebuilt || more.empty()) break;
    TNL_CHECK(pass < 2, "apply plans of the MPO sum did not stabilise");
  }
  if (ap->sharded)
    for (auto& pen : pens)
      if (!pen.dead) build_penalty_vector(pen, proto);
}

// one term, not sharded: vout = L * W1 * W2 * R applied to vin (absent L / R: trivial-link views, see make_plan)
void Env::run_plan(ApplyPlan& p, const double* vin, double* vout) {
  const bool cp = p.cplx;
  const int64_t n = p.nelem;                               // plane distance of the Krylov vectors
  const double* vin_i = cp ? vin + n : nullptr;
  double* vout_i = cp ? vout + n : nullptr;
  auto imc = [&](const TensorP& t) -> const double* { return t->cplx ? t->im() : nullptr; };
  auto imw = [&](const TensorP& t) -> double* { return t->cplx ? t->im() : nullptr; };
  auto xf = [&](TransformPlan& x, const double* sr, const double* si, double* dr, double* di, const double* W) {
    run_transform(ctx, x, sr, dr, W);
    if (cp) run_transform(ctx, x, si, di, W);
  };
  if (p.nsite == 0) {
    if (p.L && p.R) {
      cgemm(ctx, *p.g1, p.L->d, imc(p.L), false, vin, vin_i, false, p.T1->d, imw(p.T1));
      xf(*p.x2, p.T1->d, imc(p.T1), p.T3->d, imw(p.T3), nullptr);
      cgemm(ctx, *p.g4, p.T3->d, imc(p.T3), false, p.R->d, imc(p.R), false, vout, vout_i);
    } else if (p.L) {
      cgemm(ctx, *p.g1, p.L->d, imc(p.L), false, vin, vin_i, false, vout, vout_i);
    } else {
      cgemm(ctx, *p.g4, vin, vin_i, false, p.R->d, imc(p.R), false, vout, vout_i);
    }
    return;
  }
  const double* t1 = vin;
  const double* t1i = vin_i;
  if (p.g1) {
    cgemm(ctx, *p.g1, p.L->d, imc(p.L), false, vin, vin_i, false, p.T1->d, imw(p.T1));
    t1 = p.T1->d;
    t1i = imc(p.T1);
  }
  if (p.nsite == 2) {
    xf(*p.x2, t1, t1i, p.T2->d, imw(p.T2), p.W1->d);
    xf(*p.x3, p.T2->d, imc(p.T2), p.T3->d, imw(p.T3), p.W2->d);
  } else {
    xf(*p.x2, t1, t1i, p.T3->d, imw(p.T3), p.W1->d);
  }
  if (p.g4) cgemm(ctx, *p.g4, p.T3->d, imc(p.T3), false, p.R->d, imc(p.R), false, p.P->d, imw(p.P));
  xf(*p.x5, p.P->d, imc(p.P), vout, vout_i, nullptr);
}

void Env::apply_ptr(const Tensor& proto, const double* vin, double* vout) {
  if (!parent) ensure_plan(proto);
  ApplyPlan& p = *ap;
  if (p.sharded) {
    op_to_local(vin, p.loc_in);
    apply_local(p.loc_in, p.loc_out);
    op_gather(p.loc_out, vout);
    return;
  }
  if (cm) cm_apply(proto, vin, vout);
  else run_plan(p, vin, vout);
  // + sum_k H_k v   (ProjMPOSum2.product, src/mps/projmposum2.jl:86-101)
  if (!more.empty()) {
    const int64_t nv = proto.planes() * proto.nelem;
    double* tmp = ctx->vec_acquire((size_t)nv);
    for (auto& m : more) {
      m->apply_ptr(proto, vin, tmp);
      vec_axpy(ctx, vout, tmp, nv, 1.0);
    }
    ctx->vec_release(tmp, 0);
  }
  // + weight * sum_M <m|v> |m>   (ProjMPO_MPS2.product)
  for (auto& pen : pens) {
    if (pen.dead) continue;
    build_penalty_vector(pen, proto);
    if (proto.cplx) {
      vec_cdot(ctx, pen.m->d, vin, proto.nelem, 200);           // <m|v> = sum conj(m) v
      vec_caxpy_dev(ctx, vout, pen.m->d, proto.nelem, 200, weight);
    } else {
      vec_dot(ctx, pen.m->d, vin, proto.nelem, 200);
      vec_axpy_dev(ctx, vout, pen.m->d, proto.nelem, 200, weight);
    }
  }
  if (!parent) ctx->cnt.apply_count += 1;
}

double Env::expectation(const Tensor& phi) {
  Tensor tmp(ctx, phi.inds, 1, true, phi.cplx);
  apply(phi, tmp);
  vec_dot(ctx, phi.d, tmp.d, phi.planes() * phi.nelem, 0);     // real part of <phi|H|phi>: one flat pass over both planes
  fetch_scalars(ctx, 1);
  return ctx->h_scalars[0];
}

// ------------------------------------------------------------------------------------ Krylov solvers
// (templates in krylov.hpp; the sharded Krylov vectors are r-slices, one per rank)
int64_t Env::op_nloc() const { return ap->nloc; }
bool Env::op_sharded() const { return ap && ap->sharded; }
// full Krylov vector <-> this rank's r-slice (planar complex: plane by plane; full planes at distance nelem, local
// planes at distance nloc)
void Env::op_to_local(const double* full, double* loc) {
  for (int pl = 0; pl < (ap->cplx ? 2 : 1); pl++)
    run_transform(ctx, *ap->xs, full + pl * ap->nelem, loc + pl * ap->nloc, nullptr);
}
void Env::op_gather(const double* loc, double* full) {
  const int64_t nl = ap->nloc, W = ctx->world;
  for (int pl = 0; pl < (ap->cplx ? 2 : 1); pl++) {
    double* pk = ap->packed + pl * W * nl;
    comm_allgather(ctx, loc + pl * nl, pk, nl);
    for (int k = 0; k < W; k++) run_transform(ctx, *ap->unpack[k], pk + (int64_t)k * nl, full + pl * ap->nelem, nullptr);
  }
}
LanczosResult Env::eigsolve(Tensor& phi, double tol, int krylovdim, int maxiter, bool eager) {
  return krylov_eigsolve(ctx, *this, phi, tol, krylovdim, maxiter, eager);
}
ExpResult Env::exponentiate(Tensor& phi, double t_re, double t_im, double tol, int krylovdim, int maxiter, bool eager) {
  return krylov_exponentiate(ctx, *this, phi, t_re, t_im, tol, krylovdim, maxiter, eager);
}

}  // namespace tnl

namespace tnl {

// ------------------------------------------------------------------------------- noise term
// ITensorMPS `noiseterm(::ProjMPO, phi, ortho)`: nt = L*W_j*phi (left) | phi*W_{j+1}*R (right); the density
// perturbation nt*dag(noprime(nt)) is formed inside factorize() as X X^T / X^T X per charge group.
TensorP Env::noise_tensor(const Tensor& phi, bool left, bool own_storage) {
  TNL_CHECK(nsite == 2 && phi.rank() == 4, "noise term only defined for 2-site ProjMPO");
  const int first = lpos + 1;
  return left ? noise_operand(phi, true, lproj().get(), *Wlr[first - 1], own_storage)
              : noise_operand(phi, false, rproj().get(), *Wnr[first], own_storage);
}

// One noise operand X with drho = X X^T (left) / X^T X (right), for one term of the Hamiltonian:
//   left : X[(l' s1') | (w s2 r)] = E(l',w,l) W(w; s1', s1; w') phi      E = left environment or absent
//   right: X[(l s1 w) | (s2' r')] = phi W(w; s2', s2; w') E(r', w', r)  E = right environment or absent
// (ITensorMPS noiseterm(::ProjMPO); per id in src/mps/projcouplingmodel.jl:391-492).  The first operand of a bond
// may live in workspace slot T2; further operands (MPO sums, CouplingModel ids) own their memory.
TensorP Env::noise_operand(const Tensor& phi, bool left, const Tensor* E, const Tensor& W, bool own_storage) {
  const bool cx = phi.cplx || (E && E->cplx);      // ComplexF64: planar complex GEMMs, the (real) W transform per plane
  auto mk_x = [&](std::vector<Index> inds, int nrow) {
    return own_storage ? mk(ctx, std::move(inds), nrow, cx) : mk_ws(ctx, std::move(inds), nrow, Ctx::SLOT_T2, cx);
  };
  ap.reset();                                     // reuse the apply workspaces (the plan dies with the bond anyway)
  ctx->slot_epoch++;
  if (left) {
    TensorP T1;
    if (E) {
      T1 = mk_ws(ctx, {E->inds[0], E->inds[1], phi.inds[1], phi.inds[2], phi.inds[3]}, 2, Ctx::SLOT_T1, cx);
      auto g1 = plan_gemm(*E, false, phi, false, *T1);
      cgemm(ctx, *g1, *E, false, phi, false, *T1);
    } else {
      TNL_CHECK(is_trivial_link(W.inds[0]), "term without a left environment must start with a trivial link");
      T1 = view_with_trivial(ctx, phi, 1, 2, trivial_like(W.inds[0], -W.inds[0].dir), phi.d);
    }
    auto X = mk_x({T1->inds[0], W.inds[2], W.inds[3], phi.inds[2], phi.inds[3]}, 2);
    auto x = plan_transform(*T1, *X, {0, -1, -1, 3, 4}, &W, {1, 2});
    TNL_CHECK(T1->cplx == X->cplx, "noise operand element type");
    run_transform_c(ctx, *x, *T1, *X, W.d);
    ctx->sync();
    return X;
  }
  TensorP P3 = relayout(ctx, phi, 3);                           // [(l s1 s2) | r]
  TensorP Xa;
  Index rout;
  if (E) {
    TensorP R = relayout(ctx, *E, 2);                           // [(r' w) | r]
    Xa = mk_ws(ctx, {phi.inds[0], phi.inds[1], phi.inds[2], R->inds[0], R->inds[1]}, 3, Ctx::SLOT_T1, cx);
    auto g1 = plan_gemm(*P3, false, *R, true, *Xa);
    cgemm(ctx, *g1, *P3, false, *R, false, *Xa);
    ctx->sync();
    rout = R->inds[0];
  } else {
    TNL_CHECK(is_trivial_link(W.inds[1]), "term without a right environment must end with a trivial link");
    Xa = view_with_trivial(ctx, *P3, 4, 3, trivial_like(W.inds[1], -W.inds[1].dir), P3->d);
    rout = phi.inds[3];
  }
  auto X = mk_x({phi.inds[0], phi.inds[1], W.inds[2], W.inds[3], rout}, 3);       // W = (s, wr | wl, s')
  auto x = plan_transform(*Xa, *X, {0, 1, -1, -1, 3}, &W, {2, 4});
  TNL_CHECK(Xa->cplx == X->cplx, "noise operand element type");
  run_transform_c(ctx, *x, *Xa, *X, W.d);
  ctx->sync();
  return X;
}

// ITensorMPS `replacebond!(psi, pos, phi; ortho, maxdim, mindim, cutoff, eigen_perturbation, normalize)`
FactorizeResult Env::replacebond(int pos, const Tensor& phi, FactorizeParams prm, bool normalize) {
  TNL_CHECK(pos >= 1 && pos < N && phi.rank() == 4, "replacebond: bad bond / tensor");
  TNL_CHECK(!parent, "replacebond is an operation of the top environment");
  TensorP X;
  std::vector<TensorP> Xmore;
  if (prm.noise != 0.0) {
    if (cm) {
      Xmore = cm_noise_operands(phi, prm.ortho_left != 0);
      TNL_CHECK(!Xmore.empty(), "CouplingModel noise term: no term touches the kept site");
      X = Xmore.back();
      Xmore.pop_back();
      for (auto& x : Xmore) prm.noiseXmore.push_back(x.get());
      prm.noiseX = X.get();
    }
    // noiseterm(::ProjMPOSum2) = sum of the terms' noise terms (src/mps/projmposum2.jl:130-145)
    for (auto& m : more) {
      m->nsite = nsite;
      Xmore.push_back(m->noise_tensor(phi, prm.ortho_left != 0, true));
      prm.noiseXmore.push_back(Xmore.back().get());
    }
    if (!cm) {
      X = noise_tensor(phi, prm.ortho_left != 0, false);
      prm.noiseX = X.get();
    }
  }
  ap.reset();
  ctx->slot_epoch++;
  TensorP S = relayout_ws(ctx, phi, 2, Ctx::SLOT_T3);
  FactorizeResult f = factorize(ctx, *S, prm);
  if (normalize) {
    Tensor& t = prm.ortho_left ? *f.R : *f.L;
    vec_dot(ctx, t.d, t.d, t.planes() * t.nelem, 0);
    fetch_scalars(ctx, 1);
    double nrm = std::sqrt(ctx->h_scalars[0]);
    TNL_CHECK(nrm > 0, "replacebond: zero norm");
    vec_scale(ctx, t.d, t.planes() * t.nelem, 1.0 / nrm);
  }
  A[pos - 1] = f.L;
  A[pos] = f.R;
  invalidate(pos, pos + 1);
  return f;
}

// One-site split (src/mps/update_site.jl:158-186 without the TDVP reverse step):
//   U, S, V = svd(phi, uinds; maxdim, mindim, cutoff); normalize!(S); psi[pos] = U; psi[posnext] = (S*V) * psi[posnext]
FactorizeResult Env::svd_split(int pos, const Tensor& phi, FactorizeParams prm, bool normalize, bool absorb) {
  TNL_CHECK(pos >= 1 && pos <= N && phi.rank() == 3, "svd_split: bad site / tensor");
  prm.which = 1;                                    // `svd`, not `factorize`: always the SVD path
  prm.noise = 0.0;
  const bool left = prm.ortho_left != 0;
  TNL_CHECK(left ? pos < N : pos > 1, "svd_split: no neighbour in that direction");
  TensorP T = relayout(ctx, phi, left ? 2 : 1);
  FactorizeResult f = factorize(ctx, *T, prm);
  Tensor& carry = left ? *f.R : *f.L;               // S*V (m, r)  or  U*S (l, m)
  if (normalize) {
    vec_dot(ctx, carry.d, carry.d, carry.planes() * carry.nelem, 0);
    fetch_scalars(ctx, 1);
    double nrm = std::sqrt(ctx->h_scalars[0]);
    TNL_CHECK(nrm > 0, "svd_split: zero norm");
    vec_scale(ctx, carry.d, carry.planes() * carry.nelem, 1.0 / nrm);
  }
  A[pos - 1] = left ? f.L : f.R;
  invalidate(pos, pos);
  if (absorb) absorb_bond(pos, left, left ? *f.R : *f.L);
  return f;
}

// psi[posnext] = carry * psi[posnext]  (src/mps/update_site.jl:186); carry = (m, r) [left] or (l, m) [right]
void Env::absorb_bond(int pos, bool left, const Tensor& carry) {
  TNL_CHECK(carry.rank() == 2, "absorb: the bond tensor must have two indices");
  if (left) {
    TNL_CHECK(pos < N, "absorb: no right neighbour");
    TensorP C1 = carry.nrow == 1 ? nullptr : relayout(ctx, carry, 1);
    const Tensor& Cm = C1 ? *C1 : carry;
    TensorP nx = as_nrow(ctx, A[pos], 1);
    auto An = std::make_shared<Tensor>(ctx, std::vector<Index>{Cm.inds[0], nx->inds[1], nx->inds[2]}, 1, true, Cm.cplx || nx->cplx);
    auto g = plan_gemm(Cm, false, *nx, false, *An);
    cgemm(ctx, *g, Cm, false, *nx, false, *An);
    ctx->sync();
    A[pos] = An;
    invalidate(pos + 1, pos + 1);
  } else {
    TNL_CHECK(pos > 1, "absorb: no left neighbour");
    TensorP pv = as_nrow(ctx, A[pos - 2], 2);
    TensorP C1 = carry.nrow == 1 ? nullptr : relayout(ctx, carry, 1);
    const Tensor& Cm = C1 ? *C1 : carry;
    auto An = std::make_shared<Tensor>(ctx, std::vector<Index>{pv->inds[0], pv->inds[1], Cm.inds[1]}, 2, true, pv->cplx || Cm.cplx);
    auto g = plan_gemm(*pv, false, Cm, false, *An);
    cgemm(ctx, *g, *pv, false, Cm, false, *An);
    ctx->sync();
    A[pos - 2] = An;
    invalidate(pos - 1, pos - 1);
  }
}

// ITensorMPS `orthogonalize!`: QR gauge moves of the centre from site `from` to site `to` (no truncation).
// move_center(N, 1) right-canonicalises an arbitrary MPS (sweep.jl:100-102).
void Env::move_center(int from, int to) {
  TNL_CHECK(from >= 1 && from <= N && to >= 1 && to <= N, "site out of range");
  // ComplexF64: the gauge move comes from the Hermitian eigenproblem (factorize_complex, which == 3)
  FactorizeParams prm;
  prm.which = 3;
  for (int j = from; j > to; j--) {                // right-orthonormalise site j, push the rest into j-1
    TensorP Aq = as_nrow(ctx, A[j - 1], 1);        // [l | s r]
    prm.ortho_left = 0;
    FactorizeResult f = factorize(ctx, *Aq, prm);  // L (l, m) ; R = Q (m, s, r)
    A[j - 1] = f.R;
    TensorP Ap = as_nrow(ctx, A[j - 2], 2);        // [(l0 s0) | l]
    TensorP Cm = as_nrow(ctx, f.L, 1);
    auto An = std::make_shared<Tensor>(ctx, std::vector<Index>{Ap->inds[0], Ap->inds[1], f.L->inds[1]}, 2, true,
                                       Ap->cplx || Cm->cplx);
    auto g = plan_gemm(*Ap, false, *Cm, false, *An);
    cgemm(ctx, *g, *Ap, false, *Cm, false, *An);
    ctx->sync();
    A[j - 2] = An;
  }
  for (int j = from; j < to; j++) {                // left-orthonormalise site j, push the rest into j+1
    TensorP As = as_nrow(ctx, A[j - 1], 2);        // [l s | r]
    prm.ortho_left = 1;
    FactorizeResult f = factorize(ctx, *As, prm);  // L = Q (l, s, m) ; R (m, r)
    A[j - 1] = f.L;
    TensorP An1 = as_nrow(ctx, A[j], 1);           // [r | s2 r2]
    auto An = std::make_shared<Tensor>(ctx, std::vector<Index>{f.R->inds[0], An1->inds[1], An1->inds[2]}, 1, true,
                                       f.R->cplx || An1->cplx);
    auto g = plan_gemm(*f.R, false, *An1, false, *An);
    cgemm(ctx, *g, *f.R, false, *An1, false, *An);
    ctx->sync();
    A[j] = An;
  }
  invalidate(std::min(from, to), std::max(from, to));
}

}  // namespace tnl

// =================================================================================================
// ProjCouplingModel (src/mps/projcouplingmodel.jl): the Hamiltonian is a set of terms ("ids"), each present on a
// subset of the sites.  On the device every id is an MPO-like term over its support: site operators in the
// (wl, s', s, wr) form with TRIVIAL (dim 1, charge 0) links where the reference tensor has no OpLink, identity
// fillers on sites the term skips, and environments E_id(l', w, l).  Environments whose link is trivial (the
// reference's order-2 results, :170-189) are summed into one local tensor per bond.
// =================================================================================================
namespace tnl {

struct Env::CM {
  struct Op { TensorP Wlr, Wrl, Wnr; Index wl, wr; bool open_l = false, open_r = false; };
  struct E { TensorP t; bool open = false; };          // environment of one id; open: carries a real OpLink
  std::vector<std::map<int64_t, Op>> M;               // CouplingModel.terms: per site, id -> operator
  std::vector<std::map<int64_t, E>> LR;               // environments per site: id -> E_id  (L: nrow 2, R: nrow 1)
  std::map<std::string, Op> fillers;                  // identity operators of pass-through sites
  int64_t next_local = -1;                            // ids of the accumulated local tensors (gen_rand_id upstream)
  std::vector<std::pair<int64_t, std::shared_ptr<ApplyPlan>>> plans;
  int plan_lpos = -1, plan_rpos = -1;
  explicit CM(int n) : M(n), LR(n) {}
};

void Env::CMDeleter::operator()(CM* p) const { delete p; }

static Env::CM::Op make_op(Ctx* ctx, const HostBlocks& hb) {
  Env::CM::Op op;
  auto lr = permute_host(hb, {0, 2, 1, 3});     // (wl, s | s', wr)
  auto rl = permute_host(hb, {1, 3, 0, 2});     // (s', wr | wl, s)
  auto nr = permute_host(hb, {2, 3, 0, 1});     // (s, wr | wl, s')
  op.Wlr = import_tensor(ctx, lr.hb, 4);
  op.Wrl = import_tensor(ctx, rl.hb, 4);
  op.Wnr = import_tensor(ctx, nr.hb, 4);
  op.wl = hb.inds[0];
  op.wr = hb.inds[3];
  return op;
}

void Env::cm_set_term(int site, int64_t id, const HostBlocks& hb, bool open_l, bool open_r) {
  TNL_CHECK(!parent && more.empty(), "a CouplingModel environment cannot be mixed with MPO terms");
  TNL_CHECK(site >= 1 && site <= N, "site out of range");
  TNL_CHECK(hb.rank == 4, "CouplingModel site tensor must be given as (wl, s', s, wr) with trivial links where it has no OpLink");
  TNL_CHECK(id >= 0, "term ids must be non-negative");
  if (!cm) cm.reset(new CM(N));
  TNL_CHECK(open_l || is_trivial_link(hb.inds[0]), "a missing left OpLink must be given as a dim-1 charge-0 index");
  TNL_CHECK(open_r || is_trivial_link(hb.inds[3]), "a missing right OpLink must be given as a dim-1 charge-0 index");
  CM::Op op = make_op(ctx, hb);
  op.open_l = open_l;
  op.open_r = open_r;
  cm->M[site - 1][id] = std::move(op);
  lpos = 0; rpos = N + 1; ap.reset();
  cm->plans.clear();
}

// identity on the site and on the link: W(wl, s', s, wr) = delta(wl, wr) delta(s', s)
static const Env::CM::Op& cm_filler(Ctx* ctx, Env::CM& cm, const Index& site_ket, const Index& wl, const Index& wr) {
  std::string key;
  auto add = [&](const Index& ix) {
    key += std::to_string(ix.dir) + ":";
    for (int k = 0; k < ix.nsect(); k++) {
      key += std::to_string(ix.dims[k]) + "/";
      for (int a = 0; a < ix.nq; a++) key += std::to_string(ix.qns[k][a]) + ",";
    }
    key += "|";
  };
  add(site_ket); add(wl); add(wr);
  auto it = cm.fillers.find(key);
  if (it != cm.fillers.end()) return it->second;
  TNL_CHECK(wl.dims == wr.dims && wl.qns == wr.qns && wl.dir == -wr.dir, "pass-through links must be a dagged pair");
  HostBlocks hb;
  hb.rank = 4;
  Index sp = site_ket, sk = site_ket;
  sk.dir = -site_ket.dir;                       // (s' : arrow of the ket index, s : its dagger) as in an MPO tensor
  hb.inds = {wl, sp, sk, wr};
  std::vector<double> data;
  for (int a = 0; a < wl.nsect(); a++)
    for (int m = 0; m < site_ket.nsect(); m++) {
      const int dw = wl.dims[a], ds = site_ket.dims[m];
      hb.coords.push_back({a, m, m, a});
      hb.offsets.push_back((int64_t)data.size());
      const size_t base = data.size();
      data.resize(base + (size_t)dw * ds * ds * dw, 0.0);
      for (int w = 0; w < dw; w++)
        for (int x = 0; x < ds; x++)
          data[base + w + (size_t)dw * (x + (size_t)ds * (x + (size_t)ds * w))] = 1.0;
    }
  hb.data = data.data();
  return cm.fillers.emplace(key, make_op(ctx, hb)).first->second;
}

static bool same_space_dag(const Index& a, const Index& b) { return a.dims == b.dims && a.qns == b.qns && a.dir == -b.dir; }

// local_tensor += phidag (src/mps/projcouplingmodel.jl:170-189): same layout for every contribution; a real
// accumulator meeting a complex contribution is promoted first
static void add_local(Ctx* ctx, TensorP& local, const TensorP& t) {
  if (!local) { local = t; return; }
  TNL_CHECK(local->nelem == t->nelem, "local environment layouts differ");
  if (t->cplx && !local->cplx) {
    auto n = std::make_shared<Tensor>(ctx, local->inds, local->nrow, true, true);
    vec_copy(ctx, n->d, local->d, local->nelem);
    local = n;
  }
  vec_axpy(ctx, local->d, t->d, t->planes() * t->nelem, 1.0);
  ctx->sync();
}

// ProjCouplingModel._makeL! (src/mps/projcouplingmodel.jl:123-196)
void Env::cm_makeL(int k) {
  int ll = lpos;
  if (ll >= k) { lpos = k; return; }
  ll = std::max(ll, 0);
  while (ll < k) {
    const std::map<int64_t, CM::E> empty;
    const auto& L = ll <= 0 ? empty : cm->LR[ll - 1];
    const auto& Ms = cm->M[ll];
    std::map<int64_t, CM::E> next;
    TensorP local;
    std::vector<int64_t> ids;
    for (auto& kv : L) ids.push_back(kv.first);
    for (auto& kv : Ms) if (!L.count(kv.first)) ids.push_back(kv.first);
    const Index& sket = A[ll]->inds[1];
    for (int64_t id : ids) {
      auto li = L.find(id);
      auto mi = Ms.find(id);
      const Tensor* Lid = li == L.end() ? nullptr : li->second.t.get();
      const CM::Op* op;
      bool open;
      if (mi != Ms.end()) {
        op = &mi->second;
        if (Lid) TNL_CHECK(same_space_dag(Lid->inds[1], op->wl), "CouplingModel: OpLink of a term does not match its left environment");
        TNL_CHECK((Lid && li->second.open) == op->open_l, "CouplingModel: a term's left OpLink has no partner");
        open = op->open_r;
      } else {
        op = &cm_filler(ctx, *cm, sket, with_dir(Lid->inds[1], -Lid->inds[1].dir), Lid->inds[1]);
        open = li->second.open;
      }
      TensorP Ln = step_left(Lid, A[ll], *op->Wlr);
      if (!open) {
        TNL_CHECK(is_trivial_link(Ln->inds[1]), "CouplingModel: closed term with a non-trivial link");
        add_local(ctx, local, Ln);
      } else {
        next[id] = CM::E{Ln, true};
      }
    }
    if (local) next[cm->next_local--] = CM::E{local, false};
    cm->LR[ll] = std::move(next);
    ll++;
  }
  lpos = k;
}

// ProjCouplingModel._makeR! (src/mps/projcouplingmodel.jl:212-286)
void Env::cm_makeR(int k) {
  int rl = rpos;
  if (rl <= k) { rpos = k; return; }
  rl = std::min(rl, N + 1);
  while (rl > k) {
    const int j = rl - 2;                              // 0-based site being absorbed
    const std::map<int64_t, CM::E> empty;
    const auto& R = rl >= N + 1 ? empty : cm->LR[rl - 1];
    const auto& Ms = cm->M[j];
    std::map<int64_t, CM::E> next;
    TensorP local;
    std::vector<int64_t> ids;
    for (auto& kv : R) ids.push_back(kv.first);
    for (auto& kv : Ms) if (!R.count(kv.first)) ids.push_back(kv.first);
    const Index& sket = A[j]->inds[1];
    for (int64_t id : ids) {
      auto ri = R.find(id);
      auto mi = Ms.find(id);
      const Tensor* Rid = ri == R.end() ? nullptr : ri->second.t.get();
      const CM::Op* op;
      bool open;
      if (mi != Ms.end()) {
        op = &mi->second;
        if (Rid) TNL_CHECK(same_space_dag(Rid->inds[1], op->wr), "CouplingModel: OpLink of a term does not match its right environment");
        TNL_CHECK((Rid && ri->second.open) == op->open_r, "CouplingModel: a term's right OpLink has no partner");
        open = op->open_l;
      } else {
        op = &cm_filler(ctx, *cm, sket, Rid->inds[1], with_dir(Rid->inds[1], -Rid->inds[1].dir));
        open = ri->second.open;
      }
      TensorP Rn = step_right(Rid, A[j], *op->Wrl);
      if (!open) {
        TNL_CHECK(is_trivial_link(Rn->inds[1]), "CouplingModel: closed term with a non-trivial link");
        add_local(ctx, local, Rn);
      } else {
        next[id] = CM::E{Rn, true};
      }
    }
    if (local) next[cm->next_local--] = CM::E{local, false};
    cm->LR[j] = std::move(next);
    rl--;
  }
  rpos = k;
}

// the pieces of one id at the current position: environment, site operators (fillers where the id skips a site)
struct CMPieces { TensorP L, R; const Env::CM::Op* W[2] = {nullptr, nullptr}; };

static std::vector<std::pair<int64_t, CMPieces>> cm_collect(Env& e, Env::CM& cm) {
  const std::map<int64_t, Env::CM::E> empty;
  const auto& L = e.lpos <= 0 ? empty : cm.LR[e.lpos - 1];
  const auto& R = e.rpos >= e.N + 1 ? empty : cm.LR[e.rpos - 1];
  std::vector<int64_t> ids;
  auto push = [&](int64_t id) { if (std::find(ids.begin(), ids.end(), id) == ids.end()) ids.push_back(id); };
  for (auto& kv : L) push(kv.first);
  for (int sidx = 0; sidx < e.nsite; sidx++)
    for (auto& kv : cm.M[e.lpos + sidx]) push(kv.first);
  for (auto& kv : R) push(kv.first);
  std::vector<std::pair<int64_t, CMPieces>> out;
  for (int64_t id : ids) {
    CMPieces pc;
    auto li = L.find(id);
    auto ri = R.find(id);
    if (li != L.end()) pc.L = li->second.t;
    if (ri != R.end()) pc.R = ri->second.t;
    Index cur = pc.L ? pc.L->inds[1] : Index{};        // link carried to the right (as it sits on the environment)
    bool have = (bool)pc.L;
    bool open = pc.L && li->second.open;                // the carried link is a real OpLink
    for (int sidx = 0; sidx < e.nsite; sidx++) {
      const int site = e.lpos + sidx;                    // 0-based
      auto mi = cm.M[site].find(id);
      if (mi != cm.M[site].end()) {
        pc.W[sidx] = &mi->second;
        if (have) TNL_CHECK(same_space_dag(cur, mi->second.wl), "CouplingModel: OpLinks of a term do not chain");
        TNL_CHECK(open == mi->second.open_l, "CouplingModel: a term's OpLink has no partner on the left");
        open = mi->second.open_r;
      } else {
        Index t;
        if (!have) {                                     // nothing to the left: trivial pass-through
          t.nq = e.A[site]->inds[1].nq; t.dir = -1; t.dims = {1}; t.qns = {charge_zero()};
          cur = t;
        }
        pc.W[sidx] = &cm_filler(e.ctx, cm, e.A[site]->inds[1], with_dir(cur, -cur.dir), cur);
      }
      cur = pc.W[sidx]->wr;
      have = true;
    }
    if (pc.R) {
      TNL_CHECK(!have || same_space_dag(cur, pc.R->inds[1]), "CouplingModel: OpLink does not match the right environment");
      TNL_CHECK(open == ri->second.open, "CouplingModel: a term's OpLink has no partner on the right");
    } else {
      TNL_CHECK(!open, "CouplingModel: a term ends with an open OpLink");
    }
    out.emplace_back(id, pc);
  }
  return out;
}

void Env::cm_ensure_plans(const Tensor& proto) {
  TNL_CHECK(rpos - lpos == nsite + 1, "environments are not positioned for this nsite");
  for (int pass = 0; pass < 4; pass++) {
    bool stale = !ap || cm->plans.empty() || cm->plan_lpos != lpos || cm->plan_rpos != rpos || ap->nsite != nsite ||
                 ap->nelem != proto.nelem || ap->cplx != proto.cplx;
    for (auto& pl : cm->plans) stale = stale || pl.second->epoch != ctx->slot_epoch;
    if (!stale) return;
    TNL_CHECK(pass < 3, "apply plans of the CouplingModel did not stabilise");
    cm->plans.clear();
    for (auto& kv : cm_collect(*this, *cm)) {
      const CMPieces& pc = kv.second;
      cm->plans.emplace_back(kv.first, make_plan(proto, pc.L, pc.W[0] ? pc.W[0]->Wlr : nullptr,
                                                 pc.W[1] ? pc.W[1]->Wlr : nullptr, pc.R, false));
    }
    TNL_CHECK(!cm->plans.empty(), "CouplingModel has no term at this position");
    cm->plan_lpos = lpos; cm->plan_rpos = rpos;
    ap = cm->plans[0].second;                            // representative (nsite / nelem / epoch bookkeeping)
    const uint64_t e = ctx->slot_epoch;                  // a slot grew while the later plans were built?
    bool ok = true;
    for (auto& pl : cm->plans) ok = ok && pl.second->epoch == e;
    if (ok) return;
  }
}

// ProjCouplingModel.product (src/mps/projcouplingmodel.jl:315-383): sum over ids of contract(v, tensors of the id)
// Multi-GPU: the sum over term ids (src/mps/projcouplingmodel.jl:340-352) is the natural axis -- every rank applies
// the ids it owns (longest-processing-time first on the plans' flops, the same assignment on every rank) to the whole
// replicated vector and the partial results are summed with one all-reduce.
static std::vector<char> cm_my_plans(Ctx* ctx, size_t nplans, const std::function<double(size_t)>& flops) {
  const int W = ctx->shard_world();
  std::vector<char> mine(nplans, 1);
  if (W <= 1 || nplans < 2) return mine;
  std::vector<size_t> ord(nplans);
  std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return flops(a) > flops(b); });
  std::vector<double> load(W, 0.0);
  for (size_t i : ord) {
    const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    load[r] += flops(i);
    mine[i] = r == ctx->rank;
  }
  return mine;
}

void Env::cm_apply(const Tensor& proto, const double* vin, double* vout) {
  const int64_t nv = proto.planes() * proto.nelem;
  const std::vector<char> mine = cm_my_plans(ctx, cm->plans.size(), [&](size_t i) { return cm->plans[i].second->flops; });
  const bool spread = ctx->shard_world() > 1 && cm->plans.size() >= 2;
  double* tmp = nullptr;
  bool first = true;
  for (size_t i = 0; i < cm->plans.size(); i++) {
    if (!mine[i]) continue;
    auto& pl = cm->plans[i];
    if (first) { run_plan(*pl.second, vin, vout); first = false; continue; }
    if (!tmp) tmp = ctx->vec_acquire((size_t)nv);
    run_plan(*pl.second, vin, tmp);
    vec_axpy(ctx, vout, tmp, nv, 1.0);
  }
  if (tmp) ctx->vec_release(tmp, 0);
  if (spread) {
    if (first) CUDA_OK(cudaMemsetAsync(vout, 0, (size_t)nv * sizeof(double), ctx->stream));   // this rank owns no id here
    comm_allreduce_sum(ctx, vout, nv);
  }
}

double Env::cm_apply_flops() const {
  const std::vector<char> mine = cm_my_plans(ctx, cm->plans.size(), [&](size_t i) { return cm->plans[i].second->flops; });
  double f = 0.0;
  for (size_t i = 0; i < cm->plans.size(); i++) if (mine[i]) f += cm->plans[i].second->flops;
  return f;
}

// noiseterm(::ProjCouplingModel) (src/mps/projcouplingmodel.jl:391-492): one operand per id that touches the
// environment or the site operator on the kept side
std::vector<TensorP> Env::cm_noise_operands(const Tensor& phi, bool left) {
  TNL_CHECK(nsite == 2 && phi.rank() == 4, "noise term only defined for 2-site ProjMPO");
  std::vector<TensorP> out;
  for (auto& kv : cm_collect(*this, *cm)) {
    const CMPieces& pc = kv.second;
    const int site = left ? lpos : lpos + 1;
    const bool on_site = cm->M[site].count(kv.first) > 0;
    if (left ? !(pc.L || on_site) : !(pc.R || on_site)) continue;
    out.push_back(left ? noise_operand(phi, true, pc.L.get(), *pc.W[0]->Wlr, true)
                       : noise_operand(phi, false, pc.R.get(), *pc.W[1]->Wnr, true));
  }
  return out;
}

bool Env::complex_at_position() const {
  auto cx = [](const TensorP& t) { return t && t->cplx; };
  bool any = false;
  if (cm) {
    if (lpos >= 1) for (auto& kv : cm->LR[lpos - 1]) any = any || cx(kv.second.t);
    if (rpos <= N) for (auto& kv : cm->LR[rpos - 1]) any = any || cx(kv.second.t);
  } else {
    if (lpos >= 1) any = any || cx(LR[lpos - 1]);
    if (rpos <= N) any = any || cx(LR[rpos - 1]);
  }
  for (auto& m : more) any = any || m->complex_at_position();
  return any;
}

}  // namespace tnl
