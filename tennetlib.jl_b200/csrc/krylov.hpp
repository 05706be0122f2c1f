// Krylov solvers of the hot path as templates over the operator: the MPS effective Hamiltonian (Env) and the generic
// sum-of-products operator of the tree tensor networks (SumOp) run the SAME Lanczos `eigsolve` / `exponentiate`
// restatements (reference call sites: src/base/solver.jl:23-88, reached from src/mps/update_site.jl:48,83,115,182 and
// src/ttn/update_site_ttn.jl:60).  An operator provides
//   ensure_plan(proto), op_sharded(), op_nloc(), apply_local(vin, vout), apply_ptr(proto, vin, vout),
//   op_to_local(full, loc), op_gather(loc, full).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <numeric>

#include "env.hpp"

namespace tnl {

// ------------------------------------------------------------------------------------ Lanczos
// small dense symmetric eigenproblem (cyclic Jacobi); eigenvalues ascending, eigenvectors in columns
inline void sym_eig(int n, std::vector<double>& Amat, std::vector<double>& D, std::vector<double>& U) {
  U.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) U[i * n + i] = 1.0;
  auto a = [&](int i, int j) -> double& { return Amat[(size_t)i * n + j]; };
  auto u = [&](int i, int j) -> double& { return U[(size_t)i * n + j]; };
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, dia = 0;
    for (int i = 0; i < n; i++) {
      dia += a(i, i) * a(i, i);
      for (int j = i + 1; j < n; j++) off += a(i, j) * a(i, j);
    }
    if (off <= 1e-32 * dia || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a(p, q);
        if (apq == 0.0) continue;
        double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = a(k, p), akq = a(k, q);
          a(k, p) = c * akp - s * akq;
          a(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = a(p, k), aqk = a(q, k);
          a(p, k) = c * apk - s * aqk;
          a(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double ukp = u(k, p), ukq = u(k, q);
          u(k, p) = c * ukp - s * ukq;
          u(k, q) = s * ukp + c * ukq;
        }
      }
  }
  std::vector<int> perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return a(x, x) < a(y, y); });
  D.resize(n);
  std::vector<double> U2((size_t)n * n);
  for (int j = 0; j < n; j++) {
    D[j] = a(perm[j], perm[j]);
    // sign convention: largest-magnitude component positive (makes results reproducible across backends)
    int im = 0;
    for (int i = 1; i < n; i++)
      if (std::fabs(u(i, perm[j])) > std::fabs(u(im, perm[j]))) im = i;
    double sg = u(im, perm[j]) < 0 ? -1.0 : 1.0;
    for (int i = 0; i < n; i++) U2[(size_t)i * n + j] = sg * u(i, perm[j]);
  }
  U.swap(U2);
}

// KrylovKit `_householder!` for a real vector: H = I - beta v v^T, H x = nu e_i, nu >= 0
inline void householder(const std::vector<double>& x, int i, double& beta, std::vector<double>& v, double& nu) {
  v = x;
  double sigma = 0;
  for (int k = 0; k < (int)x.size(); k++)
    if (k != i) sigma += x[k] * x[k];
  double vi = x[i];
  nu = std::sqrt(vi * vi + sigma);
  if (sigma == 0.0 && vi == nu) { beta = 0.0; return; }
  if (vi < 0) vi = vi - nu; else vi = -sigma / (vi + nu);
  for (auto& e : v) e /= vi;
  v[i] = 1.0;
  beta = -vi / nu;
}

// Krylov vectors of one solver call: acquired from the context's pool, released when the call ends -- also when it
// ends with an exception (RAII)
struct VecPool {
  Ctx* c;
  size_t n;
  std::vector<double*> owned;
  VecPool(Ctx* ctx, size_t len) : c(ctx), n(len) {}
  double* get() { double* p = c->vec_acquire(n); owned.push_back(p); return p; }      // zero-filled
  void give(double* p) {
    auto it = std::find(owned.begin(), owned.end(), p);
    if (it == owned.end()) return;
    owned.erase(it);
    c->vec_release(p, 0);
  }
  ~VecPool() { for (double* p : owned) c->vec_release(p, 0); }
};

// Restatement of KrylovKit.eigsolve(A, x0, 1, :SR, Lanczos(krylovdim, maxiter, tol, eager; orth = MGS2)).
// The same algorithm is restated independently in oracle/krylov.py.  Float64 and ComplexF64 (planar) vectors: H_eff is
// Hermitian, so the tridiagonal matrix (alpha, beta) is real in both cases and only the MGS coefficients are complex;
// they stay on the device as (re, im) scalar slot pairs.  Every MGS step is ONE fused pass (vec_axpy_dot).
template <class Op>
LanczosResult krylov_eigsolve(Ctx* ctx, Op& op, Tensor& phi, double tol, int krylovdim, int maxiter, bool eager) {
  TNL_CHECK(krylovdim >= 1 && krylovdim < LC_MAX_HOST, "krylovdim out of range (1 .. 31)");
  const bool cx = phi.cplx;
  op.ensure_plan(phi);
  const bool sh = op.op_sharded();                       // Krylov vectors live as r-slices, one per rank
  const int64_t n = sh ? op.op_nloc() : phi.nelem;       // plane size
  const int64_t nv = cx ? 2 * n : n;                 // doubles per Krylov vector
  const int howmany = 1;
  std::vector<double*> V;                            // Krylov basis (device buffers)
  // pads between charge groups must stay zero: the flat BLAS-1 kernels run over the padded buffer
  VecPool pool(ctx, (size_t)nv);
  auto applyraw = [&](double* vin, double* vout) { if (sh) op.apply_local(vin, vout); else op.apply_ptr(phi, vin, vout); };
  // scalar slot pair k = (2k, 2k+1) = (re, im); inner products are reduced over the ranks on the stream
  auto reduce = [&](int k) { if (sh) comm_allreduce_sum(ctx, ctx->d_scalars + 2 * k, cx ? 2 : 1); };
  auto dot = [&](const double* x, const double* y, int k) {
    if (cx) vec_cdot(ctx, x, y, n, 2 * k); else tnl::vec_dot(ctx, x, y, n, 2 * k);
    reduce(k);
  };
  // fused MGS step: w += a * (kin >= 0 ? s[kin] : 1) * x ; s[kout] = <y, w>
  auto step = [&](double* w, const double* x, int kin, double a, const double* y, int kout) {
    if (cx) vec_caxpy_cdot(ctx, w, x, n, kin >= 0 ? 2 * kin : -1, a, 0.0, y, 2 * kout);
    else vec_axpy_dot(ctx, w, x, n, kin >= 0 ? 2 * kin : -1, a, y, 2 * kout);
    reduce(kout);
  };
  double* x0 = phi.d;
  double* phi_loc = nullptr;
  if (sh) {
    phi_loc = pool.get();
    op.op_to_local(phi.d, phi_loc);
    x0 = phi_loc;
  }
  // ---- initialize
  double* r = pool.get();
  dot(x0, x0, 0);
  fetch_scalars(ctx, 1);
  const double beta0 = std::sqrt(ctx->h_scalars[0]);
  TNL_CHECK(beta0 > 0, "initial vector should not have norm zero");
  applyraw(x0, r);                                      // A x0
  dot(x0, r, 0);
  fetch_scalars(ctx, 1);
  double alpha = ctx->h_scalars[0] / (beta0 * beta0);
  double* v0 = pool.get();
  vec_scale_to(ctx, v0, x0, nv, 1.0 / beta0);
  vec_scale(ctx, r, nv, 1.0 / beta0);
  step(r, v0, -1, -alpha, v0, 0);                       // r -= alpha v0 ; s0 = <v0, r>   (MGS2 correction)
  step(r, v0, 0, -1.0, r, 1);                           // r -= s0 v0    ; s1 = |r|^2
  fetch_scalars(ctx, 4);
  alpha += ctx->h_scalars[0];
  double beta = std::sqrt(ctx->h_scalars[2]);
  V.push_back(v0);
  std::vector<double> alphas{alpha}, betas{beta};
  LanczosResult res;
  res.numops = 1;
  res.numiter = 1;
  int converged = 0;
  std::vector<double> D, U, f;
  int Klast = 0;
  while (true) {
    beta = betas.back();
    const int K = (int)alphas.size();
    if (K == krylovdim || beta <= tol || (eager && K >= howmany)) {
      Klast = K;
      if (K == 1) {
        D = {alphas[0]}; U = {1.0}; f = {beta};
        converged = beta <= tol ? 1 : 0;
      } else {
        std::vector<double> T((size_t)K * K, 0.0);
        for (int j = 0; j < K; j++) T[(size_t)j * K + j] = alphas[j];
        for (int j = 0; j + 1 < K; j++) T[(size_t)j * K + j + 1] = T[(size_t)(j + 1) * K + j] = betas[j];
        sym_eig(K, T, D, U);
        f.resize(K);
        for (int j = 0; j < K; j++) f[j] = U[(size_t)(K - 1) * K + j] * beta;
        converged = 0;
        while (converged < K && std::fabs(f[converged]) <= tol) converged++;
      }
      if (converged >= howmany) break;
    }
    if (K < krylovdim) {
      // expand! + lanczosrecurrence (ModifiedGramSchmidt2)
      const double bold = betas.back();
      double* vnew = r;
      vec_scale(ctx, vnew, nv, 1.0 / bold);
      V.push_back(vnew);
      double* w = pool.get();
      applyraw(vnew, w);
      res.numops++;
      const int m = (int)V.size();
      step(w, V[m - 2], -1, -bold, vnew, 0);            // w -= beta v_prev ; s0 = <vnew, w>
      const double* prev = vnew;
      int slot = 0;
      for (int q = 0; q < m; q++) {                     // w -= s prev ; s_{1+q} = <V[q], w>
        step(w, prev, slot, -1.0, V[q], 1 + q);
        prev = V[q];
        slot = 1 + q;
      }
      step(w, prev, slot, -1.0, w, 1 + m);              // last correction ; |w|^2
      fetch_scalars(ctx, 2 * (2 + m));
      const double a = ctx->h_scalars[0] + ctx->h_scalars[2 * m];     // alpha + last correction (against vnew)
      const double b = std::sqrt(ctx->h_scalars[2 * (1 + m)]);
      alphas.push_back(a);
      betas.push_back(b);
      r = w;
    } else {
      if (res.numiter == maxiter) break;
      const int keep = (3 * krylovdim + 2 * converged) / 5;
      // restore tridiagonal form of [diag(D[:keep]); f[:keep]^T] with Householder reflections
      std::vector<double> H((size_t)(keep + 1) * keep, 0.0);
      auto h = [&](int i, int j) -> double& { return H[(size_t)i * keep + j]; };
      for (int j = 0; j < keep; j++) { h(j, j) = D[j]; h(keep, j) = f[j]; }
      std::vector<double> U2 = U;      // K x K row-major
      for (int j = keep - 1; j >= 0; j--) {
        std::vector<double> x(j + 1), hv;
        for (int c = 0; c <= j; c++) x[c] = h(j + 1, c);
        double hb, nu;
        householder(x, j, hb, hv, nu);
        h(j + 1, j) = nu;
        for (int c = 0; c < j; c++) h(j + 1, c) = 0.0;
        if (hb != 0.0) {
          for (int c = 0; c < keep; c++) {            // rows 0..j from the left
            double sacc = 0;
            for (int i = 0; i <= j; i++) sacc += hv[i] * h(i, c);
            for (int i = 0; i <= j; i++) h(i, c) -= hb * hv[i] * sacc;
          }
          for (int i = 0; i <= j; i++) {              // columns 0..j of rows 0..j from the right
            double sacc = 0;
            for (int c = 0; c <= j; c++) sacc += h(i, c) * hv[c];
            for (int c = 0; c <= j; c++) h(i, c) -= hb * sacc * hv[c];
          }
          for (int i = 0; i < K; i++) {               // accumulate into U
            double sacc = 0;
            for (int c = 0; c <= j; c++) sacc += U2[(size_t)i * K + c] * hv[c];
            for (int c = 0; c <= j; c++) U2[(size_t)i * K + c] -= hb * sacc * hv[c];
          }
        }
      }
      // basistransform!: B_new[j] = sum_i B[i] U2[i, j]   (real coefficients: one flat pass over both planes)
      std::vector<double*> newV;
      for (int j = 0; j < keep; j++) {
        double* y = pool.get();
        std::vector<double> coef(K);
        for (int i = 0; i < K; i++) coef[i] = U2[(size_t)i * K + j];
        vec_lincomb(ctx, y, V.data(), coef.data(), K, nv);
        newV.push_back(y);
      }
      for (double* p : V) pool.give(p);
      V = newV;
      alphas.resize(keep);
      betas.resize(keep);
      for (int j = 0; j < keep; j++) { alphas[j] = h(j, j); betas[j] = h(j + 1, j); }
      vec_scale(ctx, r, nv, betas.back() / beta);      // B[keep+1] = r/beta ; shrink!: r <- that * normres
      res.numiter++;
    }
  }
  // eigenvector = B * U[:, 0]
  {
    const int K = Klast;
    std::vector<double> coef(K);
    for (int i = 0; i < K; i++) coef[i] = U[(size_t)i * K + 0];
    TNL_CHECK((int)V.size() >= K, "Krylov basis bookkeeping");
    vec_lincomb(ctx, x0, V.data(), coef.data(), K, nv);
  }
  if (sh) op.op_gather(phi_loc, phi.d);
  ctx->sync();
  res.eval = D[0];
  res.converged = converged;
  res.normres = std::fabs(f[0]);
  return res;
}

// ------------------------------------------------------------------------------ exponentiate
// phi_1(z) = (e^z - 1)/z and phi_2(z) = (e^z - 1 - z)/z^2, series for small |z|
inline std::complex<double> phi_fn(std::complex<double> z, int order) {
  if (std::abs(z) < 0.5) {
    std::complex<double> term = order == 1 ? 1.0 : 0.5, sum = 0.0;
    for (int k = 0; k < 30; k++) { sum += term; term *= z / double(k + order + 1); }
    return sum;
  }
  const std::complex<double> e = std::exp(z);
  return order == 1 ? (e - 1.0) / z : (e - 1.0 - z) / (z * z);
}

// Restatement of KrylovKit.exponentiate(A, t, x0; Lanczos(krylovdim, maxiter, tol, eager)) = expintegrator with
// p = 1 (reference call site src/base/solver.jl:66-88): u(t) = u0 + t phi_1(tA) A u0.  One extra apply w1 = A u0,
// Lanczos factorisation started from w1 (MGS2, same recurrence as eigsolve), columns K+1 / K+2 of the exponential
// of the augmented (K+2)x(K+2) matrix = phi_1(s dt T) e1 / phi_2(s dt T) e1 (evaluated through the eigensystem of
// the real tridiagonal T on the host), error estimate |dt beta normres expH[K,K+2]| against eta = tol/|t| per
// unit time, adaptive sub-steps (gamma = 0.8) when the basis is full, eager exit at every K, first-correction
// term.  The same algorithm is restated independently in oracle/krylov.py.  All vectors stay in HBM.
template <class Op>
ExpResult krylov_exponentiate(Ctx* ctx, Op& op, Tensor& phi, double t_re, double t_im, double tol, int krylovdim, int maxiter,
                              bool eager) {
  using cd = std::complex<double>;
  const bool cx = phi.cplx;
  TNL_CHECK(t_im == 0.0 || cx, "a complex time step needs a complex (planar) vector: promote phi first");
  TNL_CHECK(krylovdim >= 1 && krylovdim < LC_MAX_HOST, "krylovdim out of range");
  ExpResult res;
  const cd t(t_re, t_im);
  const double tau = std::abs(t);
  if (tau == 0.0) { res.converged = 1; return res; }
  op.ensure_plan(phi);
  const bool sh = op.op_sharded();
  const int64_t n = sh ? op.op_nloc() : phi.nelem;        // plane size
  const int64_t nv = cx ? 2 * n : n;                  // doubles per Krylov vector
  const cd sgn = t / tau;
  VecPool pool(ctx, (size_t)nv);
  auto newvec = [&]() { return pool.get(); };
  auto applyraw = [&](double* vin, double* vout) { if (sh) op.apply_local(vin, vout); else op.apply_ptr(phi, vin, vout); res.numops++; };
  // fused MGS step: w += a * (kin >= 0 ? s[kin] : 1) * x ; s[kout] = <y, w>   (slot pairs (2k, 2k+1))
  auto step = [&](double* w, const double* x, int kin, double a, const double* y, int kout) {
    if (cx) vec_caxpy_cdot(ctx, w, x, n, kin >= 0 ? 2 * kin : -1, a, 0.0, y, 2 * kout);
    else vec_axpy_dot(ctx, w, x, n, kin >= 0 ? 2 * kin : -1, a, y, 2 * kout);
    if (sh) comm_allreduce_sum(ctx, ctx->d_scalars + 2 * kout, cx ? 2 : 1);
  };
  // <x, y> into scalar slots (2k, 2k+1) = (re, im); the imaginary part of a real product is left untouched
  auto dot = [&](const double* x, const double* y, int k) {
    if (cx) vec_cdot(ctx, x, y, n, 2 * k); else tnl::vec_dot(ctx, x, y, n, 2 * k);
    if (sh) comm_allreduce_sum(ctx, ctx->d_scalars + 2 * k, cx ? 2 : 1);
  };
  auto norm2 = [&](const double* x, int k) {          // |x|^2: one flat pass over both planes
    tnl::vec_dot(ctx, x, x, nv, 2 * k);
    if (sh) comm_allreduce_sum(ctx, ctx->d_scalars + 2 * k, 1);
  };
  double* w0 = phi.d;
  double* phi_loc = nullptr;
  if (sh) {
    phi_loc = newvec();
    op.op_to_local(phi.d, phi_loc);
    w0 = phi_loc;
  }
  std::vector<double*> V;
  std::vector<double> alphas, betas;
  double* r = nullptr;
  double* w1 = newvec();
  double beta = 0.0;
  auto release_basis = [&]() {
    for (double* p : V) pool.give(p);
    V.clear();
    if (r) pool.give(r);
    r = nullptr;
  };
  // LanczosIterator initialize on x = w1 (not consumed)
  auto lanczos_init = [&]() {
    release_basis();
    alphas.clear(); betas.clear();
    r = newvec();
    applyraw(w1, r);
    dot(w1, r, 0);
    fetch_scalars(ctx, 2);
    double alpha = ctx->h_scalars[0] / (beta * beta);
    double* v0 = newvec();
    vec_scale_to(ctx, v0, w1, nv, 1.0 / beta);
    vec_scale(ctx, r, nv, 1.0 / beta);
    step(r, v0, -1, -alpha, v0, 0);
    step(r, v0, 0, -1.0, r, 1);
    fetch_scalars(ctx, 4);
    alphas.push_back(alpha + ctx->h_scalars[0]);
    betas.push_back(std::sqrt(ctx->h_scalars[2]));
    V.push_back(v0);
  };
  auto start = [&]() -> bool {          // w1 = A w0, beta = |w1|; false: w0 is a fixed point
    applyraw(w0, w1);
    norm2(w1, 0);
    fetch_scalars(ctx, 1);
    beta = std::sqrt(ctx->h_scalars[0]);
    return beta >= tol;
  };
  // small exponential: c1 = phi_1(s dt T) e1, c2last = [phi_2(s dt T) e1]_K ; returns the error estimate
  std::vector<cd> c1;
  cd c2last = 0.0;
  auto small_exp = [&](double dt) {
    const int K = (int)alphas.size();
    std::vector<double> T((size_t)K * K, 0.0), D, Q;
    for (int j = 0; j < K; j++) T[(size_t)j * K + j] = alphas[j];
    for (int j = 0; j + 1 < K; j++) T[(size_t)j * K + j + 1] = T[(size_t)(j + 1) * K + j] = betas[j];
    if (K == 1) { D = {alphas[0]}; Q = {1.0}; } else sym_eig(K, T, D, Q);
    c1.assign(K, cd(0.0));
    c2last = 0.0;
    for (int j = 0; j < K; j++) {
      const cd p1 = phi_fn(sgn * dt * D[j], 1), p2 = phi_fn(sgn * dt * D[j], 2);
      const double q0 = Q[(size_t)0 * K + j];
      for (int i = 0; i < K; i++) c1[i] += Q[(size_t)i * K + j] * q0 * p1;
      c2last += Q[(size_t)(K - 1) * K + j] * q0 * p2;
    }
    return std::abs(dt * beta * betas.back() * c2last);
  };
  // w0 += beta * s * dt * (V c1 + c2last * r)
  auto take = [&](double dt) {
    const int K = (int)alphas.size();
    std::vector<const double*> xs(V.begin(), V.begin() + K);
    xs.push_back(r);
    std::vector<double> cr, ci;
    for (int i = 0; i < K; i++) { cr.push_back(c1[i].real()); ci.push_back(c1[i].imag()); }
    cr.push_back(c2last.real()); ci.push_back(c2last.imag());
    double* y = newvec();
    const cd f = beta * sgn * dt;
    if (cx) {
      vec_clincomb(ctx, y, xs.data(), cr.data(), ci.data(), K + 1, n);
      vec_caxpy(ctx, w0, y, n, f.real(), f.imag());
    } else {
      vec_lincomb(ctx, y, xs.data(), cr.data(), K + 1, n);
      vec_axpy(ctx, w0, y, n, f.real());
    }
    pool.give(y);
  };
  const double gamma = 0.8;
  const double eta = tol / tau;
  double tau0 = 0.0, dtau = tau, totalerr = 0.0;
  res.numiter = 1;
  bool fixed_point = !start();
  if (!fixed_point) lanczos_init();
  while (!fixed_point) {
    const int K = (int)alphas.size();
    if (K == krylovdim) {
      dtau = std::min(dtau, tau - tau0);
      double eps = small_exp(dtau);
      double omega = eps / (dtau * eta);
      double q = K / 2.0;
      while (omega > 1.0) {
        const double eps_prev = eps, dtau_prev = dtau;
        dtau *= std::pow(gamma / omega, 1.0 / (q + 1.0));
        eps = small_exp(dtau);
        omega = eps / (dtau * eta);
        q = std::max(0.0, std::log(eps / eps_prev) / std::log(dtau / dtau_prev) - 1.0);
      }
      totalerr += eps;
      take(dtau);
      tau0 += dtau;
      if (omega < gamma) dtau *= omega > 0.0 ? std::pow(gamma / omega, 1.0 / (q + 1.0)) : 1.2;
    } else if (betas.back() <= (tau - tau0) * eta || eager) {
      const double eps = small_exp(tau - tau0);
      const double omega = eps / ((tau - tau0) * eta);
      if (omega < 1.0) {
        totalerr += eps;
        take(tau - tau0);
        tau0 = tau;
      }
    }
    if (tau0 >= tau) { res.converged = 1; break; }
    if (K < krylovdim) {
      // expand! + lanczosrecurrence (ModifiedGramSchmidt2)
      const double bold = betas.back();
      double* vnew = r;
      vec_scale(ctx, vnew, nv, 1.0 / bold);
      V.push_back(vnew);
      double* w = newvec();
      r = nullptr;
      applyraw(vnew, w);
      const int m = (int)V.size();
      step(w, V[m - 2], -1, -bold, vnew, 0);
      const double* prev = vnew;
      int slot = 0;
      for (int qi = 0; qi < m; qi++) {
        step(w, prev, slot, -1.0, V[qi], 1 + qi);
        prev = V[qi];
        slot = 1 + qi;
      }
      step(w, prev, slot, -1.0, w, 1 + m);
      fetch_scalars(ctx, 2 * (2 + m));
      alphas.push_back(ctx->h_scalars[0] + ctx->h_scalars[2 * m]);       // alpha + last correction (against vnew)
      betas.push_back(std::sqrt(ctx->h_scalars[2 * (1 + m)]));
      r = w;
    } else {
      if (res.numiter == maxiter) { res.converged = 0; break; }
      res.numiter++;
      if (!start()) { fixed_point = true; break; }
      lanczos_init();
    }
  }
  if (fixed_point) { res.converged = 1; totalerr = beta; }
  release_basis();
  if (sh) op.op_gather(phi_loc, phi.d);
  ctx->sync();
  res.err = totalerr;
  return res;
}


}  // namespace tnl
