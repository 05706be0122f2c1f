// Generic block-sparse tensor algebra on the device (see algebra.hpp for the reference call sites it serves).
#include "algebra.hpp"

#include <algorithm>
#include <numeric>

namespace tnl {

// =================================================================================================
// index permutation between two charge-fused layouts
// =================================================================================================
// One CTA per 32 x 32 tile of the two indices that exchange the stride-1 role: Y's leading index (stride 1 in Y) and
// X's leading index (stride 1 in X, sitting at position `a` of Y).  The tile goes through shared memory like a dense
// transpose, so both the read and the write are coalesced; the remaining indices are enumerated by the grid.  When
// both tensors share the leading index the tile is a straight copy over (index 0, index b).
struct PermBlock {
  int64_t xoff, yoff;
  int d[MAXR];               // dims, Y order
  int64_t xst[MAXR];         // X strides of Y's indices
  int64_t yst[MAXR];
  int r;
  int a;                     // Y position of X's leading index (0: same leading index)
  int b;                     // second tile index in Y (= a, or 1 when a == 0; -1 for rank 1)
  int nt0, nt1;
  int64_t tile_start;
};

__global__ void __launch_bounds__(256)
permute_kernel(const PermBlock* __restrict__ blocks, int nblocks, const double* __restrict__ X, double* __restrict__ Y,
               int64_t ntiles) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int64_t tid = blockIdx.x; tid < ntiles; tid += gridDim.x) {
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (blocks[mid].tile_start <= tid) lo = mid; else hi = mid - 1;
    }
    const PermBlock& pb = blocks[lo];
    int64_t t = tid - pb.tile_start;
    const int t0 = (int)(t % pb.nt0); t /= pb.nt0;
    const int t1 = (int)(t % pb.nt1); t /= pb.nt1;
    int64_t xbase = pb.xoff, ybase = pb.yoff;
    for (int k = 1; k < pb.r; k++) {
      if (k == pb.b) continue;
      const int i = (int)(t % pb.d[k]); t /= pb.d[k];
      xbase += i * pb.xst[k];
      ybase += i * pb.yst[k];
    }
    const int d0 = pb.d[0], db = pb.b >= 0 ? pb.d[pb.b] : 1;
    const int64_t xs0 = pb.xst[0], ys0 = pb.yst[0];
    const int64_t xsb = pb.b >= 0 ? pb.xst[pb.b] : 0, ysb = pb.b >= 0 ? pb.yst[pb.b] : 0;
    if (pb.a != 0) {
      // read with tx along X's leading index (= Y index b), write with tx along Y's leading index
      for (int j = ty; j < 32; j += 8) {
        const int i0 = t0 * 32 + j, ib = t1 * 32 + tx;
        if (i0 < d0 && ib < db) tile[j][tx] = X[xbase + i0 * xs0 + ib * xsb];
      }
      __syncthreads();
      for (int j = ty; j < 32; j += 8) {
        const int i0 = t0 * 32 + tx, ib = t1 * 32 + j;
        if (i0 < d0 && ib < db) Y[ybase + i0 * ys0 + ib * ysb] = tile[tx][j];
      }
      __syncthreads();
    } else {
      for (int j = ty; j < 32; j += 8) {
        const int i0 = t0 * 32 + tx, ib = t1 * 32 + j;
        if (i0 < d0 && ib < db) Y[ybase + i0 * ys0 + ib * ysb] = X[xbase + i0 * xs0 + ib * xsb];
      }
    }
  }
}

// finishes the tile bookkeeping of the block list and launches (X and Y base pointers of ONE plane)
static void run_perm_blocks(Ctx* ctx, std::vector<PermBlock>& pbs, const double* X, double* Y) {
  int64_t ntiles = 0;
  for (PermBlock& pb : pbs) {
    pb.nt0 = (pb.d[0] + 31) / 32;
    pb.nt1 = pb.b >= 0 ? (pb.d[pb.b] + 31) / 32 : 1;
    int64_t rest = 1;
    for (int k = 1; k < pb.r; k++)
      if (k != pb.b) rest *= pb.d[k];
    pb.tile_start = ntiles;
    ntiles += (int64_t)pb.nt0 * pb.nt1 * rest;
  }
  if (ntiles == 0) return;
  Ctx::Scope prof_scope(ctx, 1);
  PermBlock* d = ctx->upload(pbs);
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * 64);
  permute_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(d, (int)pbs.size(), X, Y, ntiles);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->free(d);
}

static PermBlock perm_block(const Block& xb, const Block& yb, const std::vector<int>& perm) {
  PermBlock pb{};
  const int r = (int)perm.size();
  pb.r = r;
  pb.xoff = xb.off;
  pb.yoff = yb.off;
  pb.a = 0;
  for (int k = 0; k < r; k++) {
    pb.d[k] = yb.d[k];
    pb.yst[k] = yb.st[k];
    pb.xst[k] = xb.st[perm[k]];
    if (perm[k] == 0) pb.a = k;
  }
  pb.b = pb.a != 0 ? pb.a : (r > 1 ? 1 : -1);
  return pb;
}

TensorP permute(Ctx* ctx, const Tensor& X, const std::vector<int>& perm, int nrow) {
  const int r = X.rank();
  TNL_CHECK((int)perm.size() == r, "permutation length differs from the tensor rank");
  std::vector<char> seen(r, 0);
  std::vector<Index> inds(r);
  for (int k = 0; k < r; k++) {
    TNL_CHECK(perm[k] >= 0 && perm[k] < r && !seen[perm[k]], "not a permutation");
    seen[perm[k]] = 1;
    inds[k] = X.inds[perm[k]];
  }
  auto Y = std::make_shared<Tensor>(ctx, inds, nrow, true, X.cplx);
  TNL_CHECK(Y->blocks.size() == X.blocks.size(), "permuted tensor has a different block set");
  std::vector<PermBlock> pbs;
  pbs.reserve(Y->blocks.size());
  for (const Block& yb : Y->blocks) {
    int xc[MAXR];
    for (int k = 0; k < r; k++) xc[perm[k]] = yb.c[k];
    const int bi = X.find(xc);
    TNL_CHECK(bi >= 0, "permute: block missing in the source");
    pbs.push_back(perm_block(X.blocks[bi], yb, perm));
  }
  run_perm_blocks(ctx, pbs, X.d, Y->d);
  if (X.cplx) run_perm_blocks(ctx, pbs, X.im(), Y->im());
  ctx->cnt.xf_bytes += 16.0 * X.planes() * (double)X.logical_elems();
  if (!X.present.empty()) {
    Y->present.resize(Y->blocks.size());
    for (size_t i = 0; i < Y->blocks.size(); i++) {
      int xc[MAXR];
      for (int k = 0; k < r; k++) xc[perm[k]] = Y->blocks[i].c[k];
      Y->present[i] = X.present[X.find(xc)];
    }
  }
  return Y;
}

// dag(X) as a tensor of its own: arrows reversed (the charge groups change their order, hence a copy), ComplexF64
// conjugated
TensorP dag_copy(Ctx* ctx, const Tensor& X) {
  const int r = X.rank();
  std::vector<Index> inds = X.inds;
  for (Index& ix : inds) ix.dir = -ix.dir;
  auto Y = std::make_shared<Tensor>(ctx, inds, X.nrow, true, X.cplx);
  std::vector<int> ident(r);
  std::iota(ident.begin(), ident.end(), 0);
  std::vector<PermBlock> pbs;
  for (const Block& yb : Y->blocks) {
    const int bi = X.find(yb.c);
    TNL_CHECK(bi >= 0, "dag: block missing in the source");
    pbs.push_back(perm_block(X.blocks[bi], yb, ident));
  }
  run_perm_blocks(ctx, pbs, X.d, Y->d);
  if (X.cplx) {
    run_perm_blocks(ctx, pbs, X.im(), Y->im());
    vec_scale(ctx, Y->im(), Y->nelem, -1.0);
  }
  Y->present = X.present;
  return Y;
}

// =================================================================================================
// contraction over shared labels
// =================================================================================================
TensorP contract(Ctx* ctx, const Tensor& A, const std::vector<int>& la, bool dagA, const Tensor& B,
                 const std::vector<int>& lb, bool dagB, std::vector<int>* lc) {
  TNL_CHECK((int)la.size() == A.rank() && (int)lb.size() == B.rank(), "one label per index");
  std::vector<int> ia, ib, fa, fb;
  for (int k = 0; k < A.rank(); k++) {
    auto it = std::find(lb.begin(), lb.end(), la[k]);
    if (it == lb.end()) { fa.push_back(k); continue; }
    ia.push_back(k);
    ib.push_back((int)(it - lb.begin()));
  }
  for (int k = 0; k < B.rank(); k++)
    if (std::find(ib.begin(), ib.end(), k) == ib.end()) fb.push_back(k);
  TNL_CHECK(!fa.empty() || !fb.empty(), "full contraction to a scalar: use the inner product");
  TNL_CHECK((int)(fa.size() + fb.size()) <= MAXR, "result rank too large");
  for (size_t k = 0; k < ia.size(); k++) {
    const Index &x = A.inds[ia[k]], &y = B.inds[ib[k]];
    TNL_CHECK(x.dims == y.dims && x.qns == y.qns, "contracted indices span different spaces");
    TNL_CHECK((dagA ? -x.dir : x.dir) == -(dagB ? -y.dir : y.dir), "contracted indices must carry opposite arrows");
  }
  // A -> [free | contracted], B -> [contracted | free]; an operand already in that form is used in place
  std::vector<int> pa = fa, pb = ib;
  pa.insert(pa.end(), ia.begin(), ia.end());
  pb.insert(pb.end(), fb.begin(), fb.end());
  auto in_place = [](const Tensor& t, const std::vector<int>& p, int nrow) {
    if (t.nrow != nrow) return false;
    for (size_t k = 0; k < p.size(); k++)
      if (p[k] != (int)k) return false;
    return true;
  };
  TensorP Ap, Bp;
  const Tensor* Au = &A;
  const Tensor* Bu = &B;
  if (!in_place(A, pa, (int)fa.size())) { Ap = permute(ctx, A, pa, (int)fa.size()); Au = Ap.get(); }
  if (!in_place(B, pb, (int)ia.size())) { Bp = permute(ctx, B, pb, (int)ia.size()); Bu = Bp.get(); }
  std::vector<Index> ci;
  std::vector<int> cl;
  for (int k : fa) { Index x = A.inds[k]; if (dagA) x.dir = -x.dir; ci.push_back(x); cl.push_back(la[k]); }
  for (int k : fb) { Index x = B.inds[k]; if (dagB) x.dir = -x.dir; ci.push_back(x); cl.push_back(lb[k]); }
  auto C = std::make_shared<Tensor>(ctx, ci, (int)fa.size(), true, A.cplx || B.cplx);
  auto g = plan_gemm(*Au, false, *Bu, false, *C, dagA, dagB);
  cgemm(ctx, *g, *Au, dagA, *Bu, dagB, *C);      // plan arrays / permuted operands are freed in stream order
  if (lc) *lc = cl;
  return C;
}

// =================================================================================================
// direct sum along one index
// =================================================================================================
TensorP directsum(Ctx* ctx, const Tensor& A, int ia, const Tensor& B, int ib) {
  const int r = A.rank();
  TNL_CHECK(B.rank() == r && ia == ib && ia >= 0 && ia < r, "directsum: same rank and the summed index at the same position");
  TNL_CHECK(A.cplx == B.cplx, "directsum: mixed element types");
  for (int k = 0; k < r; k++)
    if (k != ia) TNL_CHECK(A.inds[k].same_space(B.inds[k]) && A.inds[k].dir == B.inds[k].dir, "directsum: the other indices must be shared");
  TNL_CHECK(A.inds[ia].dir == B.inds[ib].dir, "directsum: summed indices must carry the same arrow");
  std::vector<Index> inds = A.inds;
  Index& n = inds[ia];
  const int ns_a = A.inds[ia].nsect();
  n.dims.insert(n.dims.end(), B.inds[ib].dims.begin(), B.inds[ib].dims.end());
  n.qns.insert(n.qns.end(), B.inds[ib].qns.begin(), B.inds[ib].qns.end());
  auto C = std::make_shared<Tensor>(ctx, inds, A.nrow, true, A.cplx);
  std::vector<int> ident(r);
  std::iota(ident.begin(), ident.end(), 0);
  for (int src = 0; src < 2; src++) {
    const Tensor& X = src == 0 ? A : B;
    std::vector<PermBlock> pbs;
    for (const Block& xb : X.blocks) {
      int cc[MAXR];
      for (int k = 0; k < r; k++) cc[k] = xb.c[k];
      if (src == 1) cc[ia] += ns_a;
      const int bi = C->find(cc);
      TNL_CHECK(bi >= 0, "directsum: block missing in the result");
      PermBlock pb = perm_block(xb, C->blocks[bi], ident);
      for (int k = 0; k < r; k++) pb.d[k] = xb.d[k];
      pbs.push_back(pb);
    }
    run_perm_blocks(ctx, pbs, X.d, C->d);
    if (X.cplx) run_perm_blocks(ctx, pbs, X.im(), C->im());
  }
  return C;
}

// =================================================================================================
// sum-of-products operator
// =================================================================================================
void SumOp::apply_ptr(const Tensor& proto, const double* vin, double* vout) {
  TNL_CHECK((int)vlabels.size() == proto.rank(), "SumOp: one label per index of the vector");
  Tensor v(ctx, proto.inds, proto.nrow, false, proto.cplx);
  v.d = const_cast<double*>(vin);
  v.owns = false;
  const int64_t nv = proto.planes() * proto.nelem;
  bool first = true;
  const double f0 = ctx->cnt.gemm_flops;
  for (const Term& term : terms) {
    TensorP cur;
    const Tensor* c = &v;
    std::vector<int> labels = vlabels;
    for (const Operand& op : term.ops) {
      std::vector<int> nl;
      TensorP next = contract(ctx, *c, labels, false, *op.t, op.labels, false, &nl);
      cur = next;
      c = cur.get();
      labels = nl;
    }
    TNL_CHECK(cur, "SumOp: a term without operands");
    for (int& l : labels)
      for (auto& rl : relabel)
        if (l == rl.first) { l = rl.second; break; }
    TNL_CHECK(cur->rank() == proto.rank(), "The order of the operator-vector product P*v is not equal to the order of v");
    std::vector<int> perm(proto.rank());
    bool ident = cur->nrow == proto.nrow;
    for (int k = 0; k < proto.rank(); k++) {
      auto it = std::find(labels.begin(), labels.end(), vlabels[k]);
      TNL_CHECK(it != labels.end(), "SumOp: the product lost an index of the vector");
      perm[k] = (int)(it - labels.begin());
      ident = ident && perm[k] == k;
    }
    TensorP res = ident ? cur : permute(ctx, *cur, perm, proto.nrow);
    TNL_CHECK(res->nelem == proto.nelem && res->cplx == proto.cplx, "SumOp: product layout differs from the vector layout");
    if (first) vec_copy(ctx, vout, res->d, nv);
    else vec_axpy(ctx, vout, res->d, nv, 1.0);
    first = false;
  }
  if (first) CUDA_OK(cudaMemsetAsync(vout, 0, nv * sizeof(double), ctx->stream));
  for (const TensorP& m : projs) {
    TNL_CHECK(m->nelem == proto.nelem && m->cplx == proto.cplx, "SumOp: projector layout differs from the vector layout");
    if (proto.cplx) {
      vec_cdot(ctx, m->d, vin, proto.nelem, 200);
      vec_caxpy_dev(ctx, vout, m->d, proto.nelem, 200, weight);
    } else {
      vec_dot(ctx, m->d, vin, proto.nelem, 200);
      vec_axpy_dev(ctx, vout, m->d, proto.nelem, 200, weight);
    }
  }
  flops = ctx->cnt.gemm_flops - f0;
  ctx->cnt.apply_count += 1;
}

void SumOp::apply(const Tensor& v, Tensor& out) {
  TNL_CHECK(out.nelem == v.nelem && out.nrow == v.nrow && out.cplx == v.cplx, "output vector layout mismatch");
  apply_ptr(v, v.d, out.d);
}

}  // namespace tnl
