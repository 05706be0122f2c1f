/* Drives the drop-in boundary (include/tnl_b200.h) from plain C -- no Python, no ctypes, no torch -- the way the Julia
 * shim (julia/TenNetLibB200.jl) does through `ccall`: two-site DMRG sweeps of the S=1/2 Heisenberg chain, N = 6,
 * U(1) (total Sz) block-sparse tensors in the NDTensors flat layout, Neel start, maxdim 8 (exact for N = 6).
 *
 * Call sequence per bond = the body of `_update_two_site!` (reference src/mps/update_site.jl:27-90):
 *   tnl_env_set_nsite(2) ; tnl_env_make_phi ; tnl_env_position ; tnl_eigsolve_lanczos ; tnl_vec_norm / tnl_vec_scale ;
 *   tnl_replacebond
 * Checks: ground-state energy against exact diagonalisation (dense Jacobi eigensolver below, Sz = 0 sector not
 * needed: the full 64 x 64 matrix is diagonalised), bond dimensions, export round trip of the final state's norm.
 * Built by __graft_entry__.build(); run by tests/test_c_abi.py under `-m gpu`.                                       */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tnl_b200.h"

#define N 6
#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc__ = (call);                                                               \
    if (rc__ != 0) {                                                                 \
      fprintf(stderr, "FAILED (%d) %s : %s\n", rc__, #call, tnl_last_error(ctx));    \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

static tnl_ctx_t ctx = NULL;

/* ---- a QN index with at most 8 sectors and one charge */
typedef struct { int32_t nsect, dir, dims[8], qns[8]; } idx_t;
static tnl_index_t view(const idx_t* ix) {
  tnl_index_t v; v.nsect = ix->nsect; v.dir = ix->dir; v.dims = ix->dims; v.qns = ix->qns; return v;
}
static int idx_dim(const idx_t* ix) { int d = 0; for (int s = 0; s < ix->nsect; s++) d += ix->dims[s]; return d; }
static int idx_off(const idx_t* ix, int s) { int o = 0; for (int k = 0; k < s; k++) o += ix->dims[k]; return o; }

/* flux-0 blocks of a dense column-major rank-r array -> NDTensors flat layout (coords, offsets, data) */
static int64_t dense_to_blocks(int r, const idx_t* inds, const double* dense, int32_t* coords, int64_t* offsets,
                               double* data, int64_t* nelem_out) {
  int c[4] = {0, 0, 0, 0};
  int64_t nb = 0, off = 0;
  int full[4], st[4];
  for (int k = 0; k < r; k++) full[k] = idx_dim(&inds[k]);
  st[0] = 1;
  for (int k = 1; k < r; k++) st[k] = st[k - 1] * full[k - 1];
  for (;;) {
    int q = 0;
    for (int k = 0; k < r; k++) q += inds[k].dir * inds[k].qns[c[k]];
    if (q == 0) {
      int d[4], o[4], sz = 1, nz = 0;
      for (int k = 0; k < r; k++) { d[k] = inds[k].dims[c[k]]; o[k] = idx_off(&inds[k], c[k]); sz *= d[k]; }
      int e[4] = {0, 0, 0, 0};
      for (int n = 0; n < sz; n++) {
        int src = 0;
        for (int k = 0; k < r; k++) src += (o[k] + e[k]) * st[k];
        data[off + n] = dense[src];
        if (dense[src] != 0.0) nz = 1;
        for (int k = 0; k < r; k++) { if (++e[k] < d[k]) break; e[k] = 0; }
      }
      if (nz) {                               /* ITensors drops blocks that are identically zero */
        for (int k = 0; k < r; k++) coords[nb * r + k] = c[k];
        offsets[nb] = off;
        off += sz;
        nb++;
      }
    }
    int k = 0;
    while (k < r) { if (++c[k] < inds[k].nsect) break; c[k] = 0; k++; }
    if (k == r) break;
  }
  *nelem_out = off;
  return nb;
}

/* ---- exact diagonalisation: cyclic Jacobi on the dense 2^N x 2^N Hamiltonian */
static double ed_ground_energy(void) {
  const int D = 1 << N;
  double* H = (double*)calloc((size_t)D * D, sizeof(double));
  for (int b = 0; b + 1 < N; b++)
    for (int s = 0; s < D; s++) {
      int u = (s >> b) & 1, v = (s >> (b + 1)) & 1;          /* bit = 1 : spin up */
      H[(size_t)s * D + s] += (u == v) ? 0.25 : -0.25;
      if (u != v) { int t = s ^ (1 << b) ^ (1 << (b + 1)); H[(size_t)t * D + s] += 0.5; }
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < D; p++) for (int q = p + 1; q < D; q++) off += H[(size_t)p * D + q] * H[(size_t)p * D + q];
    if (off < 1e-26) break;
    for (int p = 0; p < D - 1; p++)
      for (int q = p + 1; q < D; q++) {
        double apq = H[(size_t)p * D + q];
        if (fabs(apq) < 1e-300) continue;
        double th = (H[(size_t)q * D + q] - H[(size_t)p * D + p]) / (2 * apq);
        double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
        double cs = 1 / sqrt(t * t + 1), sn = t * cs;
        for (int k = 0; k < D; k++) {
          double akp = H[(size_t)k * D + p], akq = H[(size_t)k * D + q];
          H[(size_t)k * D + p] = cs * akp - sn * akq; H[(size_t)k * D + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < D; k++) {
          double apk = H[(size_t)p * D + k], aqk = H[(size_t)q * D + k];
          H[(size_t)p * D + k] = cs * apk - sn * aqk; H[(size_t)q * D + k] = sn * apk + cs * aqk;
        }
      }
  }
  double e = H[0];
  for (int k = 1; k < D; k++) if (H[(size_t)k * D + k] < e) e = H[(size_t)k * D + k];
  free(H);
  return e;
}

int main(void) {
  CHECK(tnl_ctx_create(0, &ctx));
  tnl_env_t env;
  CHECK(tnl_env_create(ctx, N, &env));

  /* ---- site index (Up: +1, Dn: -1 in units of 2 Sz) and the w = 5 Heisenberg MPO automaton
   * states (charge carried to the right): P "S- applied" (-2), F finished (0), Z "Sz applied" (0), I idle (0), M (+2) */
  idx_t site = {2, +1, {1, 1}, {+1, -1}};
  idx_t bulk = {3, +1, {1, 3, 1}, {-2, 0, +2}};
  idx_t edge = {1, +1, {1}, {0}};
  const double Sz[4] = {0.5, 0, 0, -0.5}, Sp[4] = {0, 0, 1, 0}, Sm[4] = {0, 1, 0, 0}, Id[4] = {1, 0, 0, 1};
  /* column-major 2x2 [s', s]: Sp = |up><dn| -> element (0,1) -> index 0 + 2*1 = 2 */
  enum { P = 0, F = 1, Z = 2, I = 3, M = 4 };
  for (int j = 0; j < N; j++) {
    idx_t inds[4];
    inds[0] = (j == 0) ? edge : bulk; inds[0].dir = +1;
    inds[1] = site; inds[1].dir = +1;
    inds[2] = site; inds[2].dir = -1;
    inds[3] = (j == N - 1) ? edge : bulk; inds[3].dir = -1;
    const int wl = idx_dim(&inds[0]), wr = idx_dim(&inds[3]);
    double* dense = (double*)calloc((size_t)wl * 4 * wr, sizeof(double));
    struct { int a, b; const double* op; double c; } tab[8] = {
        {F, F, Id, 1}, {P, F, Sp, 1}, {M, F, Sm, 1}, {Z, F, Sz, 1}, {I, P, Sm, 0.5}, {I, M, Sp, 0.5}, {I, Z, Sz, 1}, {I, I, Id, 1}};
    for (int t = 0; t < 8; t++) {
      if (j == 0 && tab[t].a != I) continue;
      if (j == N - 1 && tab[t].b != F) continue;
      const int ia = (j == 0) ? 0 : tab[t].a, ib = (j == N - 1) ? 0 : tab[t].b;
      for (int sp = 0; sp < 2; sp++)
        for (int s = 0; s < 2; s++) dense[ia + wl * (sp + 2 * (s + 2 * ib))] += tab[t].c * tab[t].op[sp + 2 * s];
    }
    int32_t coords[4 * 64]; int64_t offsets[64]; double data[256]; int64_t ne;
    int64_t nb = dense_to_blocks(4, inds, dense, coords, offsets, data, &ne);
    tnl_index_t v[4] = {view(&inds[0]), view(&inds[1]), view(&inds[2]), view(&inds[3])};
    CHECK(tnl_env_set_site_op(env, j + 1, 1, v, nb, coords, offsets, data));
    free(dense);
  }

  /* ---- Neel MPS: one 1x1x1 block per site, link charges accumulate */
  int q = 0;
  for (int j = 0; j < N; j++) {
    const int st = j % 2;                       /* 0 = Up, 1 = Dn */
    idx_t l = {1, +1, {1}, {q}};
    q += site.qns[st];
    idx_t r = {1, -1, {1}, {q}};
    idx_t inds[3] = {l, site, r};
    tnl_index_t v[3] = {view(&inds[0]), view(&inds[1]), view(&inds[2])};
    int32_t coords[3] = {0, st, 0}; int64_t off = 0; double one = 1.0;
    tnl_tensor_t A;
    CHECK(tnl_tensor_import(ctx, 3, 1, v, 1, coords, &off, &one, 2, &A));
    CHECK(tnl_env_set_state(env, j + 1, A));
    CHECK(tnl_tensor_free(A));                  /* the environment shares the tensor */
  }

  /* ---- sweeps: `fullsweep!` (src/mps/sweep.jl:118-160) with `_update_two_site!` inlined */
  double energy = 0, truncerr = 0, eigs[64];
  int64_t neigs = 0;
  for (int sweep = 0; sweep < 4; sweep++)
    for (int half = 0; half < 2; half++)
      for (int k = 0; k < N - 1; k++) {
        const int bond = half == 0 ? k + 1 : N - 1 - k;
        tnl_tensor_t phi;
        CHECK(tnl_env_set_nsite(env, 2));
        CHECK(tnl_env_make_phi(env, bond, &phi));
        CHECK(tnl_env_position(env, bond));
        int32_t conv, nops, nit; double nres, nrm;
        CHECK(tnl_eigsolve_lanczos(env, phi, 1e-14, 5, 2, 0, &energy, &conv, &nops, &nit, &nres));
        CHECK(tnl_vec_norm(phi, &nrm));
        CHECK(tnl_vec_scale(phi, 1.0 / nrm));
        CHECK(tnl_replacebond(env, bond, phi, half == 0, 8, 1, 1e-14, sweep == 0 ? 1e-3 : 0.0, 1, 0, &truncerr, eigs, 64, &neigs));
        CHECK(tnl_tensor_free(phi));
      }
  const double e_ed = ed_ground_energy();
  printf("E(device) = %.14f   E(exact) = %.14f   kept = %lld   truncerr = %.2e\n", energy, e_ed, (long long)neigs, truncerr);
  if (fabs(energy - e_ed) > 1e-10 * fabs(e_ed)) { fprintf(stderr, "energy mismatch\n"); return 2; }

  /* ---- export of the centre tensor: flat NDTensors layout back on the host, norm 1 */
  tnl_tensor_t A1;
  CHECK(tnl_env_get_state(env, 1, &A1));
  int64_t nb, ne;
  CHECK(tnl_tensor_export_size(A1, &nb, &ne));
  int32_t* coords = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(nb > 0 ? nb : 1));
  int64_t* offsets = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nb > 0 ? nb : 1));
  double* data = (double*)malloc(sizeof(double) * (size_t)(ne > 0 ? ne : 1));
  CHECK(tnl_tensor_export(A1, coords, offsets, data));
  double n2 = 0;
  for (int64_t i = 0; i < ne; i++) n2 += data[i] * data[i];
  printf("centre tensor: %lld blocks, %lld elements, norm^2 = %.14f\n", (long long)nb, (long long)ne, n2);
  if (fabs(n2 - 1.0) > 1e-12) { fprintf(stderr, "state not normalised\n"); return 3; }
  double cnt[8];
  CHECK(tnl_get_counters(ctx, cnt));
  if (cnt[4] <= 0 || cnt[5] <= 0) { fprintf(stderr, "no kernels were launched\n"); return 4; }
  free(coords); free(offsets); free(data);
  CHECK(tnl_tensor_free(A1));
  CHECK(tnl_env_destroy(env));
  CHECK(tnl_ctx_destroy(ctx));
  printf("C ABI SMOKE OK (%d kernel launches)\n", (int)cnt[4]);
  return 0;
}
